"""Sample-rate conversion on the GPU for inputs that are not at the model's rate — what FunASR's loader does with
torchaudio.transforms.Resample (funasr/utils/load_utils.py:176-178; torchaudio defaults sinc_interp_hann, lowpass_filter_width 6,
rolloff 0.99).  `sinc_resample_table` restates torchaudio.functional.functional._get_sinc_resample_kernel (including its float32
rounding of j / new_freq); the convolution itself is csrc/resample.cu (fa_resample)."""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from . import _abi


@lru_cache(maxsize=16)
def sinc_resample_table(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99) -> Tuple[np.ndarray, int, int, int]:
    """-> (table float32 [new, 2*width + orig], orig, new, width) with orig / new divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torch.arange(0, -new, -1) is int64 and `/ new_freq` yields float32 before the sum with the float64 grid promotes it
    shift = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (shift + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * window * scale
    return kern.astype(np.float32), orig, new, width


def resample(wav: torch.Tensor, lens: torch.Tensor, orig_freq: int, new_freq: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """wav [B, N] fp32 on a CUDA device, lens [B] int32 (valid samples) -> (resampled [B, ceil(new*N/orig)], new lens int32)."""
    if orig_freq == new_freq:
        return wav, lens
    if not wav.is_cuda:
        raise _abi.FunasrB200Error("funasr_b200.resample needs CUDA tensors (no CPU path)")
    table, orig, new, width = sinc_resample_table(int(orig_freq), int(new_freq))
    lib = _abi.load()
    wav = wav.to(torch.float32).contiguous()
    lens = lens.to(wav.device, torch.int32).contiguous()
    B, N = wav.shape
    cap = -(-new * N // orig)
    out = torch.empty((B, cap), dtype=torch.float32, device=wav.device)
    out_lens = torch.empty((B,), dtype=torch.int32, device=wav.device)
    tab = torch.from_numpy(table).to(wav.device)
    _abi.check(lib.fa_resample(wav.data_ptr(), lens.data_ptr(), B, wav.stride(0), tab.data_ptr(), orig, new, width, out.data_ptr(), out.stride(0),
                               cap, out_lens.data_ptr(), torch.cuda.current_stream(wav.device).cuda_stream), "fa_resample")
    return out, out_lens
