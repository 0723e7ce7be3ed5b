"""CT-Transformer punctuation restoration on the GPU behind the reference's plugin surface.

  CTTransformerB200 <- funasr/models/ct_transformer/model.py:39-485 (CTTransformer): inference(text) -> [{"key", "text", "punc_array"}]

The network (punc_forward :112-125: Embedding -> SANMEncoder d=256, 8 heads x 32, FFN 1024, 4 blocks -> Linear 256 -> |punc_list| ->
arg-max) runs as fa_embedding + fa_sanm_encoder_forward (fp32 path; the small-head attention kernel) + fa_linear_argmax.  The
text-side logic around it — word splitting (utils.py:28-99), mini-sentences of `split_size` tokens with the unfinished tail of
the previous one carried over (:330-372), capitalisation / spacing / ASCII punctuation for Latin words and the forced sentence
end (:376-441) — is host string work in the reference as well and is restated here step by step.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _abi
from .engine import _EngineBase
from .modules import SANMEncoderB200
from .registry import get_tables, register


def split_to_mini_sentence(words: list, word_limit: int = 20) -> list:
    """ct_transformer/utils.py:9-25."""
    assert word_limit > 1
    if len(words) <= word_limit:
        return [words]
    n = len(words) // word_limit
    out = [words[i * word_limit:(i + 1) * word_limit] for i in range(n)]
    if len(words) % word_limit > 0:
        out.append(words[n * word_limit:])
    return out


def split_words(text: str) -> List[str]:
    """ct_transformer/utils.py:78-99 (no jieba dictionary): ASCII runs are words, every other character is a word."""
    words: List[str] = []
    for seg in text.split():
        cur = ""
        for c in seg:
            if len(c.encode()) == 1:
                cur += c
            else:
                if cur:
                    words.append(cur)
                    cur = ""
                words.append(c)
        if cur:
            words.append(cur)
    return words


class PuncEngine(_EngineBase):
    def __init__(self, state: Dict[str, torch.Tensor], device, heads: int, ln_eps: float = 1e-12):
        self._init_base(state, device, "fp32", ln_eps)
        n = 0
        while ("encoder.encoders.%d.norm1.weight" % n) in state:
            n += 1
        names = ["encoder.encoders0.0"] + ["encoder.encoders.%d" % i for i in range(n)]
        self.d_in = int(state["embed.weight"].shape[1])
        self.d_model = int(state["encoder.after_norm.weight"].numel())
        self.enc = self._enc_stack(names, "encoder.after_norm", heads, 0, self.d_in)
        self.embed = self._g("embed.weight")
        self.out = self._lin("decoder")
        self.n_punc = int(self.out.out_f)
        torch.cuda.current_stream(self.device).synchronize()
        self._state = None

    def punc_ids(self, token_ids: np.ndarray) -> np.ndarray:
        """punc_forward + arg-max for ONE mini-sentence (batch 1, like the reference): int ids [T] -> punctuation ids [T]."""
        T = int(len(token_ids))
        ids = torch.from_numpy(np.ascontiguousarray(token_ids, dtype=np.int32)).to(self.device, non_blocking=True)
        x = torch.empty((1, T, self.d_in), dtype=torch.float32, device=self.device)
        st = self._stream()
        _abi.check(self.lib.fa_embedding(ids.data_ptr(), self.embed.data_ptr(), self.d_in, int(self.embed.shape[0]), T, x.data_ptr(), st), "fa_embedding")
        lens = torch.tensor([T], dtype=torch.int32).to(self.device, non_blocking=True)
        h = self._encode(self.enc, x, lens, self.d_model)
        out = torch.empty(T, dtype=torch.int32, device=self.device)
        best = torch.empty(T, dtype=torch.float32, device=self.device)
        ws = self._workspace(self.lib.fa_linear_argmax_workspace_bytes(T, self.n_punc, self.mode))
        _abi.check(self.lib.fa_linear_argmax(C.byref(self.out), h.data_ptr(), None, T, out.data_ptr(), best.data_ptr(), None, self.mode,
                                             ws.data_ptr(), ws.numel(), st), "fa_linear_argmax")
        return out.cpu().numpy()


@register("model_classes", "CTTransformerB200")
class CTTransformerB200(nn.Module):
    """Drop-in for CTTransformer's inference (ct_transformer/model.py:290-485)."""

    def __init__(self, encoder: str = None, encoder_conf: dict = None, vocab_size: int = -1, punc_list: list = None, punc_weight: list = None,
                 embed_unit: int = 128, att_unit: int = 256, dropout_rate: float = 0.5, ignore_id: int = -1, sos: int = 1, eos: int = 2,
                 sentence_end_id: int = 3, **kwargs):
        super().__init__()
        if kwargs.get("jieba_usr_dict") is not None:
            raise _abi.FunasrB200Error("CTTransformerB200 does not support a jieba word dictionary (word-level punctuation models)")
        conf = dict(encoder_conf or {})
        conf.setdefault("input_size", embed_unit)
        self.encoder = SANMEncoderB200(**conf)
        self.embed = nn.Embedding(vocab_size, embed_unit)
        self.decoder = nn.Linear(att_unit, len(punc_list))
        for p_ in list(self.embed.parameters()) + list(self.decoder.parameters()):
            p_.requires_grad_(False)
        self.punc_list, self.sentence_end_id = list(punc_list), sentence_end_id
        self.heads = int(conf.get("attention_heads", 8))
        self._engine: Optional[PuncEngine] = None

    def on_pretrained_model_loaded(self, loaded_keys=None):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self, device) -> PuncEngine:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _abi.FunasrB200Error("CTTransformerB200 needs a CUDA device; there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = PuncEngine(self.state_dict(), dev, self.heads)
        return self._engine

    def inference(self, data_in, data_lengths=None, key: list = None, tokenizer=None, frontend=None, **kwargs):
        assert len(data_in) == 1
        if not data_in[0] or (isinstance(data_in[0], str) and not data_in[0].strip()):
            return [{"key": key[0] if key else "", "text": "", "punc_array": None}], {"batch_data_time": -1}
        eng = self.engine(kwargs.get("device", "cuda"))
        text = data_in[0]
        split_size = kwargs.get("split_size", 20)
        tokens = split_words(text)
        tokens_int = tokenizer.encode(tokens)
        mini_sentences = split_to_mini_sentence(tokens, split_size)
        mini_sentences_id = split_to_mini_sentence(tokens_int, split_size)
        pl = self.punc_list
        cache_sent: list = []
        cache_sent_id = np.array([], dtype="int32")
        new_mini_sentence = ""
        new_mini_sentence_punc: List[int] = []
        cache_pop_trigger_limit = 200
        punc_array: Optional[np.ndarray] = None
        new_mini_sentence_out, new_mini_sentence_punc_out = "", []
        for i_s in range(len(mini_sentences)):
            mini_sentence = cache_sent + list(mini_sentences[i_s])
            mini_sentence_id = np.concatenate((cache_sent_id, np.asarray(mini_sentences_id[i_s], dtype="int32")), axis=0)
            punctuations = eng.punc_ids(mini_sentence_id).astype(np.int64)
            assert punctuations.shape[0] == len(mini_sentence)
            if i_s < len(mini_sentences) - 1:                       # carry the unfinished tail over (model.py:351-372)
                sentence_end, last_comma = -1, -1
                for i in range(len(punctuations) - 2, 1, -1):
                    if pl[punctuations[i]] == "。" or pl[punctuations[i]] == "？":
                        sentence_end = i
                        break
                    if last_comma < 0 and pl[punctuations[i]] == "，":
                        last_comma = i
                if sentence_end < 0 and len(mini_sentence) > cache_pop_trigger_limit and last_comma >= 0:
                    sentence_end = last_comma
                    punctuations[sentence_end] = self.sentence_end_id
                cache_sent = mini_sentence[sentence_end + 1:]
                cache_sent_id = mini_sentence_id[sentence_end + 1:]
                mini_sentence = mini_sentence[0:sentence_end + 1]
                punctuations = punctuations[0:sentence_end + 1]
            new_mini_sentence_punc += [int(x) for x in punctuations]
            words_with_punc = []
            for i in range(len(mini_sentence)):                     # model.py:378-403
                latin = len(mini_sentence[i][0].encode()) == 1
                if (i == 0 or pl[punctuations[i - 1]] == "。" or pl[punctuations[i - 1]] == "？") and latin:
                    mini_sentence[i] = mini_sentence[i].capitalize()
                if i == 0 and latin:
                    mini_sentence[i] = " " + mini_sentence[i]
                if i > 0 and latin and len(mini_sentence[i - 1][0].encode()) == 1:
                    mini_sentence[i] = " " + mini_sentence[i]
                words_with_punc.append(mini_sentence[i])
                if pl[punctuations[i]] != "_":
                    punc_res = pl[punctuations[i]]
                    if len(mini_sentence[i][0].encode()) == 1:
                        punc_res = {"，": ",", "。": ".", "？": "?"}.get(punc_res, punc_res)
                    words_with_punc.append(punc_res)
            new_mini_sentence += "".join(words_with_punc)
            new_mini_sentence_out, new_mini_sentence_punc_out = new_mini_sentence, new_mini_sentence_punc
            if i_s == len(mini_sentences) - 1:                      # forced sentence end (model.py:407-441)
                last = new_mini_sentence[-1] if new_mini_sentence else ""
                force = None
                if last == "，" or last == "、":
                    force = new_mini_sentence[:-1] + "。"
                elif last == ",":
                    force = new_mini_sentence[:-1] + "."
                elif last != "。" and last != "？" and last and len(last.encode()) != 1:
                    force = new_mini_sentence + "。"
                elif last != "." and last != "?" and last and len(last.encode()) == 1:
                    force = new_mini_sentence + "."
                if force is not None:
                    new_mini_sentence_out = force
                    new_mini_sentence_punc_out = new_mini_sentence_punc[:-1] + [self.sentence_end_id]
                    if len(punctuations):
                        punctuations[-1] = self.sentence_end_id
            punc_array = punctuations if punc_array is None else np.concatenate([punc_array, punctuations], axis=0)
        result = {"key": key[0] if key else "", "text": new_mini_sentence_out, "punc_array": torch.from_numpy(np.asarray(punc_array))}
        return [result], {}
