"""Flat model file for the handle-style C API (csrc/offline.cu: fa_offline_init).

One file = every tensor the path needs under FunASR's own state_dict names (so it can be produced from an unmodified
model.pt + am.mvn), plus the derived tables the kernels take as inputs:

    __config__                         [10] enc_layers, dec_layers, d_model, heads, fsmn kernel, vocab, feat_dim, ln_eps,
                                            cif threshold, tail threshold
    frontend.mel_banks [80,257], frontend.window [400], frontend.cmvn [2,560] (optional)
    encoder.pe_inv_timescales [280]    SinusoidalPositionEncoder timescales (transformer/embedding.py:409-414)
    predictor.cif_conv1d.gemm_weight   Conv1d(512,512,3) weight repacked to a [512, 3*512] GEMM weight

Layout: b"FAB2MDL1", u32 n_tensors, then per tensor: u32 name_len, name (utf-8), u32 ndim, i64 dims[ndim], u64 nbytes,
zero padding to a 16-byte file offset, little-endian fp32 data.
"""
from __future__ import annotations

import struct
from typing import Dict, Optional

import numpy as np
import torch

from .synth import ParaformerConfig, sinusoid_inv_timescales

MAGIC = b"FAB2MDL1"


def model_tensors(state: Dict[str, torch.Tensor], cfg: ParaformerConfig, cmvn: Optional[torch.Tensor]) -> Dict[str, np.ndarray]:
    from .engine import kaldi_mel_banks
    out: Dict[str, np.ndarray] = {}
    out["__config__"] = np.array([cfg.enc_layers, cfg.dec_layers, cfg.d_model, cfg.heads, cfg.kernel, cfg.vocab, cfg.feat_dim,
                                  cfg.ln_eps, cfg.cif_threshold, cfg.tail_threshold], dtype=np.float32)
    out["frontend.mel_banks"] = kaldi_mel_banks().numpy()
    out["frontend.window"] = torch.hamming_window(400, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32).numpy()
    if cmvn is not None:
        out["frontend.cmvn"] = cmvn.detach().float().cpu().numpy()
    out["encoder.pe_inv_timescales"] = sinusoid_inv_timescales(cfg.feat_dim).float().numpy()
    for k, v in state.items():
        if k.startswith(("encoder.", "predictor.", "decoder.", "bias_encoder.", "bias_embed.")) and torch.is_floating_point(v):
            out[k] = v.detach().float().cpu().contiguous().numpy()
    cw = state["predictor.cif_conv1d.weight"].detach().float().cpu()
    out["predictor.cif_conv1d.gemm_weight"] = cw.permute(0, 2, 1).reshape(cw.shape[0], -1).contiguous().numpy()
    return out


def write_model_file(path: str, state: Dict[str, torch.Tensor], cfg: ParaformerConfig, cmvn: Optional[torch.Tensor] = None) -> int:
    tensors = model_tensors(state, cfg, cmvn)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(tensors)))
        for name, arr in tensors.items():
            arr = np.ascontiguousarray(arr, dtype="<f4")
            nb = name.encode("utf-8")
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<I", arr.ndim))
            f.write(struct.pack("<%dq" % arr.ndim, *arr.shape))
            f.write(struct.pack("<Q", arr.nbytes))
            f.write(b"\0" * ((16 - f.tell() % 16) % 16))
            f.write(arr.tobytes())
    return len(tensors)


def read_model_file(path: str) -> Dict[str, np.ndarray]:
    """Python reader of the same layout (tests; mirrors load_file() in csrc/offline.cu)."""
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not a funasr_b200 model file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode("utf-8")
            (nd,) = struct.unpack("<I", f.read(4))
            shape = struct.unpack("<%dq" % nd, f.read(8 * nd)) if nd else ()
            (nbytes,) = struct.unpack("<Q", f.read(8))
            f.seek((16 - f.tell() % 16) % 16, 1)
            out[name] = np.frombuffer(f.read(nbytes), dtype="<f4").reshape(shape)
    return out
