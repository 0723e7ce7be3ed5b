"""ORACLE for the FSMN-VAD scores — CPU fp32 restatement, TEST INFRASTRUCTURE ONLY (same rules as paraformer_oracle.py).

Restates what FsmnVADStreaming feeds its end-point detector with (funasr/models/fsmn_vad_streaming):
  * frontend: kaldi Fbank (pinned to torchaudio 2.11.0, paraformer_oracle.kaldi_fbank) -> LFR m=5 n=1 -> CMVN.  The reference runs
    the stateful WavFrontendOnline over 60 s chunks (frontends/wav_frontend.py:259-660); with the (lfr_m-1)/2 replicated first
    frames and the replicated last frame on the final call that equals the clamped gather below over the whole utterance;
  * encoder FSMN.forward (encoder.py:355-377): in_linear1 -> in_linear2 -> ReLU -> 4 x [linear (no bias) -> causal depthwise
    memory (x + conv_left over lorder frames, zero left padding; FSMNBlock :136-160) -> affine -> ReLU] -> out_linear1 ->
    out_linear2 -> softmax;
  * frame energies 10 log10(sum x^2 + 1e-6) over the 400-sample frame (ComputeDecibel, model.py:458-529).
Parity status: PINNED — tests/golden/vad_*.npz hold the unmodified reference's scores and segments for the same seeded weights
and waveforms (oracle/make_vad_golden.py); tests/test_vad_host.py checks this file and funasr_b200/vad.py against them.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

import paraformer_oracle as O

Tensor = torch.Tensor


def vad_features(wav: Tensor, cmvn: Optional[Tensor], lfr_m: int = 5) -> Tensor:
    """[n] waveform in [-1, 1] -> [T, 80 * lfr_m]."""
    mel = O.kaldi_fbank(wav * (1 << 15))                                   # wav_frontend.py:420-431
    T = mel.shape[0]
    half = (lfr_m - 1) // 2
    idx = (torch.arange(T)[:, None] + torch.arange(lfr_m)[None, :] - half).clamp_(0, T - 1)
    x = mel[idx].reshape(T, -1)
    if cmvn is not None:
        x = (x + cmvn[0]) * cmvn[1]                                        # apply_cmvn :330-343
    return x.float()


def fsmn_scores(x: Tensor, p: Dict[str, Tensor], layers: int = 4) -> Tensor:
    """[T, 400] -> softmax posteriors [T, 248]."""
    def lin(name, v, bias=True):
        return F.linear(v, p[name + ".linear.weight"], p[name + ".linear.bias"] if bias else None)
    h = torch.relu(lin("encoder.in_linear2", lin("encoder.in_linear1", x)))
    for i in range(layers):
        pre = "encoder.fsmn.%d" % i
        q = lin(pre + ".linear", h, bias=False)                            # [T, 128]
        w = p[pre + ".fsmn_block.conv_left.weight"][:, 0, :, 0]            # [128, lorder]
        lo = w.shape[1]
        qp = F.pad(q.t()[None], (lo - 1, 0))                               # zero left padding, causal
        mem = F.conv1d(qp, w[:, None, :], groups=w.shape[0])[0].t()
        h = torch.relu(lin(pre + ".affine", q + mem))
    return torch.softmax(lin("encoder.out_linear2", lin("encoder.out_linear1", h)), dim=-1)


def frame_decibels(wav: Tensor) -> Tensor:
    n = (wav.numel() - 400) // 160 + 1 if wav.numel() >= 400 else 0
    if n <= 0:
        return torch.zeros(0)
    frames = wav[: (n - 1) * 160 + 400].unfold(0, 400, 160)
    return 10 * torch.log10(frames.pow(2).sum(-1) + 0.000001)


def vad_scores(wav: Tensor, p: Dict[str, Tensor], cmvn: Optional[Tensor]):
    with torch.no_grad():
        feats = vad_features(wav, cmvn)
        scores = fsmn_scores(feats, p)
    return {"feats": feats, "scores": scores, "sil_prob": scores[:, 0], "decibel": frame_decibels(wav)}
