"""FSMN-VAD on the GPU behind the reference's plugin surface (funasr/models/fsmn_vad_streaming).

  FSMNB200                  <- encoder.py:296-377 (FSMN): parameter container under the reference's names
  FsmnVADStreamingB200      <- model.py:367-1103 (FsmnVADStreaming): inference() -> [{"key", "value": [[start_ms, end_ms], ...]}]

`VadEngine` runs the fused Fbank + LFR 5/1 + CMVN kernel, the FSMN encoder (fa_fsmn_vad_forward) and the frame energies
(fa_frame_decibels) over the whole waveform, reads back two floats per 10 ms frame and hands them to the host detector
(funasr_b200/vad.py).  No torch.nn op on the path; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _abi
from .engine import FrontendEngine
from .modules import WavFrontendB200, _ParamHolder, _as_wave_list, load_cmvn
from .registry import register
from .vad import VadOptions, detect_segments_native, num_frames


def _pad16(k: int) -> int:
    return (k + 15) // 16 * 16


class VadEngine:
    """Packed FSMN weights (input dimensions zero-padded to a multiple of 16 for the fp32 GEMM) + the three kernel calls."""

    def __init__(self, state: Dict[str, torch.Tensor], device, cmvn: Optional[torch.Tensor], sil_pdf_ids=(0,), prefix: str = "encoder."):
        self.lib = _abi.load()
        self.device = torch.device(device)
        self._keep: List[torch.Tensor] = []
        self.frontend = FrontendEngine(cmvn, self.device, lfr_m=5, lfr_n=1)
        n_layers = 0
        while (prefix + "fsmn.%d.linear.linear.weight" % n_layers) in state:
            n_layers += 1
        self.layers = (_abi.FaVadLayer * max(n_layers, 1))()
        self.enc = _abi.FaVadEncoder()
        self.enc.in1 = self._lin(state, prefix + "in_linear1.linear")
        self.enc.in2 = self._lin(state, prefix + "in_linear2.linear")
        for i in range(n_layers):
            p = prefix + "fsmn.%d." % i
            self.layers[i].lin = self._lin(state, p + "linear.linear", bias=False)
            cw = state[p + "fsmn_block.conv_left.weight"]
            if (p + "fsmn_block.conv_right.weight") in state:
                raise _abi.FunasrB200Error("FSMN-VAD with a right-context memory (rorder > 0) is not supported")
            cw = cw.detach().to(self.device, torch.float32).reshape(cw.shape[0], -1).contiguous()      # [proj, lorder]
            self._keep.append(cw)
            self.layers[i].conv_w = cw.data_ptr()
            self.layers[i].affine = self._lin(state, p + "affine.linear")
            self.enc.lorder = int(cw.shape[1])
        self.enc.layers, self.enc.n_layers = self.layers, n_layers
        self.enc.out1 = self._lin(state, prefix + "out_linear1.linear")
        self.enc.out2 = self._lin(state, prefix + "out_linear2.linear")
        ids = list(sil_pdf_ids)[:4]
        for k, v in enumerate(ids):
            self.enc.sil_ids[k] = int(v)
        self.enc.n_sil = len(ids)
        self.out_dim = int(self.enc.out2.out_f)
        self._ws = None
        torch.cuda.current_stream(self.device).synchronize()

    def _lin(self, state, name, bias=True) -> _abi.FaLinear:
        w = state[name + ".weight"].detach().to(self.device, torch.float32)
        out_f, in_f = w.shape
        kp = _pad16(in_f)
        wp = torch.zeros((out_f, kp), dtype=torch.float32, device=self.device)
        wp[:, :in_f] = w
        self._keep.append(wp)
        b = None
        if bias:
            b = state[name + ".bias"].detach().to(self.device, torch.float32).contiguous()
            self._keep.append(b)
        return _abi.FaLinear(wp.data_ptr(), None if b is None else b.data_ptr(), None, out_f, kp, kp, 0)

    def scores(self, wav: torch.Tensor, want_scores: bool = False):
        """wav: 1-D fp32 on the device -> (sil_prob [T], decibel [T], scores [T, 248] or None) on the device."""
        n = int(wav.numel())
        T = num_frames(n)
        if T <= 0:
            z = torch.zeros(0, device=self.device)
            return z, z, None
        st = torch.cuda.current_stream(self.device).cuda_stream
        lens = torch.tensor([n], dtype=torch.int32).to(self.device, non_blocking=True)
        feats, _ = self.frontend(wav.reshape(1, -1), lens, T)                      # [1, T, 400]
        sil = torch.empty(T, dtype=torch.float32, device=self.device)
        db = torch.empty(T, dtype=torch.float32, device=self.device)
        sc = torch.empty((T, self.out_dim), dtype=torch.float32, device=self.device) if want_scores else None
        need = int(self.lib.fa_fsmn_vad_workspace_bytes(C.byref(self.enc), T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need * 1.1) + 1024, dtype=torch.uint8, device=self.device)
        _abi.check(self.lib.fa_fsmn_vad_forward(C.byref(self.enc), feats.data_ptr(), 400, T, sil.data_ptr(), None if sc is None else sc.data_ptr(),
                                                self._ws.data_ptr(), self._ws.numel(), st), "fa_fsmn_vad_forward")
        _abi.check(self.lib.fa_frame_decibels(wav.data_ptr(), n, T, db.data_ptr(), st), "fa_frame_decibels")
        return sil, db, sc

    def segments(self, wav: torch.Tensor, opts: Optional[VadOptions] = None, **kw) -> List[List[int]]:
        sil, db, _ = self.scores(wav)
        both = torch.stack([sil, db]).cpu().numpy() if sil.numel() else np.zeros((2, 0), np.float32)       # one D2H: two floats per frame
        # the sequential end-point walk runs in the library's host code (csrc/vad_detector.cpp); funasr_b200.vad.detect_segments is the
        # same state machine in Python, kept as the specification (tests hold the two identical)
        return detect_segments_native(both[0], both[1], int(wav.numel()), opts, **kw)


@register("encoder_classes", "FSMNB200")
class FSMNB200(_ParamHolder):
    """Parameter container for the reference's FSMN encoder (fsmn_vad_streaming/encoder.py:296-377): same constructor arguments,
    same state_dict names (in_linear1.linear.weight, fsmn.{i}.fsmn_block.conv_left.weight, ...)."""

    def __init__(self, input_dim: int, input_affine_dim: int, fsmn_layers: int, linear_dim: int, proj_dim: int, lorder: int, rorder: int,
                 lstride: int, rstride: int, output_affine_dim: int, output_dim: int, use_softmax: bool = True, **kwargs):
        super().__init__()
        if rorder != 0 or lstride != 1 or proj_dim != 128 or input_dim != 400 or lorder != 20 or not use_softmax:
            raise _abi.FunasrB200Error("FSMNB200 supports the fsmn-vad shape: input 400, proj 128, lorder 20, rorder 0, lstride 1, softmax")
        self.cfg = dict(input_dim=input_dim, input_affine_dim=input_affine_dim, fsmn_layers=fsmn_layers, linear_dim=linear_dim, proj_dim=proj_dim,
                        lorder=lorder, output_affine_dim=output_affine_dim, output_dim=output_dim)
        self._build()

    def output_size(self) -> int:
        return self.cfg["output_dim"]

    def _specs(self):
        c = self.cfg
        s = {"in_linear1.linear.weight": (c["input_affine_dim"], c["input_dim"]), "in_linear1.linear.bias": (c["input_affine_dim"],),
             "in_linear2.linear.weight": (c["linear_dim"], c["input_affine_dim"]), "in_linear2.linear.bias": (c["linear_dim"],),
             "out_linear1.linear.weight": (c["output_affine_dim"], c["linear_dim"]), "out_linear1.linear.bias": (c["output_affine_dim"],),
             "out_linear2.linear.weight": (c["output_dim"], c["output_affine_dim"]), "out_linear2.linear.bias": (c["output_dim"],)}
        for i in range(c["fsmn_layers"]):
            p = "fsmn.%d." % i
            s[p + "linear.linear.weight"] = (c["proj_dim"], c["linear_dim"])
            s[p + "fsmn_block.conv_left.weight"] = (c["proj_dim"], 1, c["lorder"], 1)
            s[p + "affine.linear.weight"] = (c["linear_dim"], c["proj_dim"])
            s[p + "affine.linear.bias"] = (c["linear_dim"],)
        return s


@register("frontend_classes", "WavFrontendOnlineB200")
class WavFrontendOnlineB200(nn.Module):
    """Stands in for WavFrontendOnline (frontends/wav_frontend.py:259-660) in a VAD config: it only carries the configuration
    (fs, LFR 5/1, CMVN); the framing itself is the fused kernel inside VadEngine, over the whole waveform."""

    def __init__(self, cmvn_file: str = None, fs: int = 16000, window: str = "hamming", n_mels: int = 80, frame_length: int = 25,
                 frame_shift: int = 10, lfr_m: int = 1, lfr_n: int = 1, dither: float = 1.0, cmvn: Optional[torch.Tensor] = None, **kwargs):
        super().__init__()
        if (fs, window, n_mels, frame_length, frame_shift, lfr_m, lfr_n) != (16000, "hamming", 80, 25, 10, 5, 1):
            raise _abi.FunasrB200Error("WavFrontendOnlineB200 is built for the fsmn-vad frontend (16 kHz, hamming 25/10 ms, 80 mel, LFR 5/1)")
        self.fs, self.frame_shift, self.lfr_m, self.lfr_n, self.cmvn_file = fs, frame_shift, lfr_m, lfr_n, cmvn_file
        self.cmvn = cmvn if cmvn is not None else (None if cmvn_file is None else load_cmvn(cmvn_file))

    def output_size(self) -> int:
        return 80 * self.lfr_m


@register("model_classes", "FsmnVADStreamingB200")
class FsmnVADStreamingB200(nn.Module):
    """Drop-in for FsmnVADStreaming's OFFLINE inference (model.py:949-1103, is_final on the last chunk): one waveform in,
    [{"key": ..., "value": [[start_ms, end_ms], ...]}] out."""

    def __init__(self, encoder: str = None, encoder_conf: Optional[dict] = None, vad_post_args=None, **kwargs):
        super().__init__()
        self.vad_opts = VadOptions.from_conf(kwargs)
        self.encoder = FSMNB200(**(encoder_conf or {}))
        self.encoder_conf = encoder_conf
        self._engine: Optional[VadEngine] = None

    def on_pretrained_model_loaded(self, loaded_keys=None):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self, device, cmvn) -> VadEngine:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _abi.FunasrB200Error("FsmnVADStreamingB200 needs a CUDA device; there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = VadEngine(self.state_dict(), dev, cmvn, self.vad_opts.sil_pdf_ids)
        return self._engine

    def inference(self, data_in, data_lengths=None, key: list = None, tokenizer=None, frontend=None, cache: dict = None, **kwargs):
        device = torch.device(kwargs.get("device", "cuda"))
        meta_data = {}
        t1 = time.perf_counter()
        wavs = _as_wave_list(data_in, fs=getattr(frontend, "fs", 16000), audio_fs=int(kwargs.get("fs", 16000)),
                             **{k: v for k, v in kwargs.items() if k not in ("fs", "audio_fs", "frontend")})
        meta_data["load_data"] = f"{time.perf_counter() - t1:0.3f}"
        if len(wavs) != 1:
            raise _abi.FunasrB200Error("batch_size must be set 1 (model.py:1001)")
        if kwargs.get("is_streaming_input", False) or not kwargs.get("is_final", True) or int(kwargs.get("chunk_size", 60000)) != 60000:
            raise _abi.FunasrB200Error("FsmnVADStreamingB200 implements the offline path (whole waveform, chunk_size 60000, is_final)")
        wav = wavs[0]
        k0 = key[0] if key else ""
        if isinstance(k0, (list, tuple)):
            k0 = k0[0]
        if wav.numel() == 0:
            return [{"key": k0, "value": []}], meta_data
        eng = self.engine(device, getattr(frontend, "cmvn", None))
        wav_dev = wav.to(device, torch.float32, non_blocking=True).contiguous()
        kw = {}
        if kwargs.get("max_end_silence_time") is not None:
            kw["max_end_silence_time"] = int(kwargs["max_end_silence_time"])
        kw["dynamic_silence"] = kwargs.get("dynamic_silence", kwargs.get("max_end_silence_time") is None)
        if kwargs.get("silence_schedule") is not None:
            kw["silence_schedule"] = kwargs["silence_schedule"]
        if kwargs.get("speech_noise_thres") is not None:
            kw["speech_noise_thres"] = float(kwargs["speech_noise_thres"])
        segments = eng.segments(wav_dev, self.vad_opts, **kw)
        meta_data["batch_data_time"] = num_frames(int(wav.numel())) * 10 / 1000
        return [{"key": k0, "value": segments}], meta_data
