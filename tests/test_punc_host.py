"""CT-Transformer host logic + oracle against the unmodified reference's goldens (no GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

import punc_oracle as PO
from funasr_b200 import synth
from funasr_b200.punc import CTTransformerB200, split_to_mini_sentence, split_words

CASES = ["punc_short", "punc_long", "punc_english_tail"]


class _CharTokenizer:
    """What the reference's CharTokenizer.encode does for a list of words: look each one up, unknown -> <unk>."""

    def __init__(self, toks):
        self.t2i = {t: i for i, t in enumerate(toks)}
        self.unk = self.t2i["<unk>"]

    def encode(self, words):
        return [self.t2i.get(w, self.unk) for w in words]


class _OracleEngine:
    def __init__(self, p):
        self.p = p

    def punc_ids(self, ids):
        return PO.punc_ids([int(i) for i in ids], self.p, synth.PUNC_LAYERS, synth.PUNC_HEADS).numpy()


def _model():
    m = CTTransformerB200(encoder="SANMEncoder",
                          encoder_conf=dict(input_size=synth.PUNC_DIM, output_size=synth.PUNC_DIM, attention_heads=synth.PUNC_HEADS,
                                            linear_units=synth.PUNC_FFN, num_blocks=synth.PUNC_LAYERS, kernel_size=11, sanm_shfit=0,
                                            input_layer="pe", normalize_before=True),
                          vocab_size=len(synth.punc_token_list()), punc_list=synth.PUNC_LIST, punc_weight=[1.0] * len(synth.PUNC_LIST),
                          embed_unit=synth.PUNC_DIM, att_unit=synth.PUNC_DIM, sentence_end_id=3)
    return m


@pytest.mark.parametrize("name", CASES)
def test_host_text_logic_with_oracle_network_equals_reference(name, monkeypatch):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = _model()
    eng = _OracleEngine(synth.make_punc_state_dict(0))
    monkeypatch.setattr(m, "engine", lambda device: eng)
    res, _ = m.inference([str(g["text_in"])], key=["k"], tokenizer=_CharTokenizer(synth.punc_token_list()))
    assert res[0]["text"] == str(g["text_out"])
    assert res[0]["punc_array"].tolist() == g["punc_array"].tolist()


def test_split_words_and_mini_sentences():
    assert split_words("gpu 你好 ok yes 世界") == ["gpu", "你", "好", "ok", "yes", "世", "界"]
    assert split_words("ab你cd") == ["ab", "你", "cd"]
    assert split_to_mini_sentence(list(range(45)), 20) == [list(range(20)), list(range(20, 40)), list(range(40, 45))]
    assert split_to_mini_sentence(list(range(20)), 20) == [list(range(20))]


def test_empty_text_returns_empty_result():
    res, _ = _model().inference([""], key=["k"], tokenizer=None)
    assert res[0]["text"] == "" and res[0]["punc_array"] is None
