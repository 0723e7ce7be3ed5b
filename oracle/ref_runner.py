"""Reference arm of bench.py (`--impl reference`) and of the AutoModel drop-in test — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (modelscope/FunASR 1.4.3) on the host CPU through its own public API:
`AutoModel(model=..., device="cpu").generate(input=...)` — funasr/auto/auto_model.py:551-829 — with the seeded synthetic
weights of funasr_b200/synth.py saved as a `model.pt` and loaded by the reference's own `load_pretrained_model`.
The reference is imported from /root/reference (build container) or from the offline install under `baseline/_ref/`
(GPU box; ref_shim.py).  If neither imports, callers fall back to the CPU restatement (oracle/paraformer_oracle.py, kind "port").

BASELINE.md §3 report: >= 5 timed runs (min / median), per-stage split, a 1-thread figure, a batch-8 run, the CPU model string.
"""
from __future__ import annotations

import os
import statistics
import sys
import tempfile
import time
from typing import Dict, List, Optional

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import ref_shim  # noqa: E402
from funasr_b200 import synth  # noqa: E402


def cpu_model_string() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def write_cmvn_file(path, cmvn):
    """Kaldi-nnet text layout parsed by load_cmvn (wav_frontend.py:15-43)."""
    d = cmvn.shape[1]
    with open(path, "w") as f:
        f.write("<Nnet>\n<Splice> %d %d\n[ 0 ]\n<AddShift> %d %d\n" % (d, d, d, d))
        f.write("<LearnRateCoef> 0 [ " + " ".join("%.9g" % v for v in cmvn[0].tolist()) + " ]\n")
        f.write("<Rescale> %d %d\n" % (d, d))
        f.write("<LearnRateCoef> 0 [ " + " ".join("%.9g" % v for v in cmvn[1].tolist()) + " ]\n</Nnet>\n")


def _enc_conf(cfg, **extra):
    d = dict(output_size=cfg.d_model, attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.enc_layers, dropout_rate=0.1,
             positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="pe", pos_enc_class="SinusoidalPositionEncoder",
             normalize_before=True, kernel_size=cfg.kernel, sanm_shfit=0, selfattention_layer_type="sanm")
    d.update(extra)
    return d


def _dec_conf(cfg):
    return dict(attention_heads=cfg.heads, linear_units=cfg.ffn, num_blocks=cfg.dec_layers, dropout_rate=0.1, positional_dropout_rate=0.1,
                self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1, att_layer_num=cfg.dec_layers, kernel_size=cfg.kernel,
                sanm_shfit=0)


def _frontend_conf(cmvn_file):
    return dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0, cmvn_file=cmvn_file)


def token_list(cfg, sv=False):
    if sv:
        return ["<blank>"] + ["t%d" % i for i in range(cfg.vocab - 2)] + ["<unk>"]
    return ["<blank>", "<s>", "</s>"] + ["t%d" % i for i in range(cfg.vocab - 4)] + ["<unk>"]


class IdTokenizer:
    """Stands in for the tokenizer so that results carry the greedy ids verbatim (SenseVoiceSmall.inference needs
    tokenizer.decode, model.py:1027)."""

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


def build_automodel(kind: str, cfg, seed: int, cmvn: Optional[torch.Tensor], tmp: str, ncpu: int, device: str = "cpu",
                    backend_keys: bool = False):
    """kind: paraformer | contextual | sensevoice.  backend_keys=True names the funasr_b200 classes' OWN keys (drop-in test);
    the default names the reference's classes (the CPU arm)."""
    ref_shim.import_reference()
    from funasr import AutoModel
    cmvn_file = None
    if cmvn is not None:
        cmvn_file = os.path.join(tmp, "am_%s.mvn" % kind)
        write_cmvn_file(cmvn_file, cmvn)
    pt = os.path.join(tmp, "%s_%d_%d.pt" % (kind, cfg.enc_layers, seed))
    common = dict(frontend="WavFrontend", frontend_conf=_frontend_conf(cmvn_file), tokenizer="CharTokenizer",
                  device=device, ncpu=ncpu, disable_update=True, disable_pbar=True, init_param=pt)
    if kind == "sensevoice":
        torch.save(synth.make_sensevoice_state_dict(cfg, seed), pt)
        return AutoModel(model="SenseVoiceSmall", model_conf=dict(length_normalized_loss=True, sos=1, eos=2, ignore_id=-1),
                         encoder="SenseVoiceEncoderSmall", encoder_conf=_enc_conf(cfg, tp_blocks=cfg.tp_layers),
                         tokenizer_conf=dict(token_list=token_list(cfg, True), unk_symbol="<unk>", split_with_space=True), **common)
    model_conf = dict(ctc_weight=0.0, lsm_weight=0.1, length_normalized_loss=True, predictor_weight=1.0, predictor_bias=1, sampling_ratio=0.75)
    pred = dict(predictor="CifPredictorV2", predictor_conf=dict(idim=cfg.d_model, threshold=1.0, l_order=1, r_order=1, tail_threshold=cfg.tail_threshold))
    tok = dict(tokenizer_conf=dict(token_list=token_list(cfg), unk_symbol="<unk>", split_with_space=True))
    if kind == "contextual":
        torch.save(synth.make_contextual_state_dict(cfg, seed), pt)
        model_conf["inner_dim"] = 512
        return AutoModel(model="ContextualParaformer", model_conf=model_conf, encoder="SANMEncoder", encoder_conf=_enc_conf(cfg),
                         decoder="ContextualParaformerDecoder", decoder_conf=_dec_conf(cfg), **pred, **tok, **common)
    torch.save(synth.make_state_dict(cfg, seed), pt)
    return AutoModel(model="Paraformer", model_conf=model_conf, encoder="SANMEncoder", encoder_conf=_enc_conf(cfg),
                     decoder="ParaformerSANMDecoder", decoder_conf=_dec_conf(cfg), **pred, **tok, **common)


def generate_ids(am, wavs: List[torch.Tensor], batch_size: int = 1, **kw) -> List[List[int]]:
    """AutoModel.generate(input=[np waveforms]) with the tokenizer removed so results carry token_int (paraformer/model.py:694)."""
    res = am.generate(input=[w.numpy() for w in wavs], batch_size=batch_size, disable_pbar=True, tokenizer=None, **kw)
    return [[int(t) for t in r["token_int"]] for r in res]


def _time_generate(am, wavs, batch_size, runs, **kw):
    ts = []
    for _ in range(runs):
        t0 = time.perf_counter()
        am.generate(input=[w.numpy() for w in wavs], batch_size=batch_size, disable_pbar=True, tokenizer=None, **kw)
        ts.append(time.perf_counter() - t0)
    return ts


def _set_ncpu(am, n: int):
    """AutoModel re-applies `ncpu` from its baseline kwargs before every generate() (auto_model.py:1335-1359)."""
    am.kwargs["ncpu"] = n
    base = getattr(am, "_base_kwargs_map", None)
    if base and isinstance(base.get("kwargs"), dict):
        base["kwargs"]["ncpu"] = n
    torch.set_num_threads(n)


def stage_split_ms(am, wav: torch.Tensor) -> Dict[str, float]:
    """Per-stage wall time of one utterance through the reference's stage methods (paraformer/model.py:286-346)."""
    from funasr.utils.load_utils import extract_fbank
    model, frontend = am.model, am.kwargs["frontend"]
    out = {}
    with torch.no_grad():
        t0 = time.perf_counter()
        feats, flens = extract_fbank([wav], frontend=frontend)
        t1 = time.perf_counter()
        enc, elens = model.encode(feats, flens)
        if isinstance(enc, tuple):
            enc = enc[0]
        t2 = time.perf_counter()
        emb, tok, _, _ = model.calc_predictor(enc, elens)[:4]
        t3 = time.perf_counter()
        model.cal_decoder_with_predictor(enc, elens, emb, tok.round().long())
        t4 = time.perf_counter()
    out.update(frontend=(t1 - t0) * 1e3, encoder=(t2 - t1) * 1e3, predictor=(t3 - t2) * 1e3, decoder=(t4 - t3) * 1e3)
    return out


def paraformer_report(am, wavs: List[torch.Tensor], threads: int, runs: int = 5, extras: bool = True) -> dict:
    """BASELINE.md §3.2-3.3 for the Paraformer arm: `runs` timed generate() calls of the sample at batch 1 (the reference's CPU
    default, auto_model.py:785), min / median; per-stage split; 1-thread figure; batch-8 figure."""
    audio = sum(w.numel() for w in wavs) / 16000.0
    ts = _time_generate(am, wavs, 1, runs)
    rep = {"runs_s": ts, "min_s": min(ts), "median_s": statistics.median(ts), "rtfx_min_time": audio / min(ts),
           "rtfx_median": audio / statistics.median(ts), "cores": threads, "cpu_model": cpu_model_string(), "torch": torch.__version__}
    if extras:
        try:
            rep["stages_ms"] = stage_split_ms(am, wavs[0])
        except Exception as e:  # pragma: no cover
            rep["stages_ms"] = {"error": repr(e)[:200]}
        w8 = [wavs[i % len(wavs)] for i in range(8)]
        t8 = _time_generate(am, w8, 8, 2)
        rep["batch8_rtfx"] = 8 * (wavs[0].numel() / 16000.0) / min(t8) if len({w.numel() for w in wavs}) == 1 else \
            sum(w.numel() for w in w8) / 16000.0 / min(t8)
        _set_ncpu(am, 1)
        try:
            t1 = _time_generate(am, wavs[:1], 1, 2)
            rep["one_thread_rtfx"] = (wavs[0].numel() / 16000.0) / min(t1)
        finally:
            _set_ncpu(am, threads)
    return rep
