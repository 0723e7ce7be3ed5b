"""ctypes binding of the handle-style C API (include/funasr_b200.h: fa_offline_*), the counterpart of FunASR's C++ runtime
FunOfflineInit / FunOfflineInferBuffer / FunASRGetResult (runtime/onnxruntime/include/funasrruntime.h:100-116).
Nothing here touches torch on the data path: host PCM buffers in, token ids out."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _abi


class OfflineRecognizer:
    def __init__(self, model_file: str, device: int = 0, gemm_mode: str = "fp16x3"):
        self.lib = _abi.load()
        mode = _abi.GEMM_MODES[gemm_mode] if isinstance(gemm_mode, str) else int(gemm_mode)
        self.handle = self.lib.fa_offline_init(model_file.encode(), device, mode)
        if not self.handle:
            raise _abi.FunasrB200Error("fa_offline_init failed: %s" % self.lib.fa_offline_last_error().decode())

    def infer(self, wavs: Sequence[np.ndarray]) -> List[List[int]]:
        """wavs: float32 arrays in [-1, 1] or int16 PCM arrays (all the same dtype), 16 kHz mono, >= 400 samples each."""
        arrs = [np.ascontiguousarray(w) for w in wavs]
        kinds = {a.dtype for a in arrs}
        if kinds == {np.dtype(np.float32)}:
            fmt = 0
        elif kinds == {np.dtype(np.int16)}:
            fmt = 1
        else:
            raise _abi.FunasrB200Error("waveforms must all be float32 or all be int16, got %s" % kinds)
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_int64 * n)(*[a.shape[0] for a in arrs])
        res = self.lib.fa_offline_infer(self.handle, ptrs, lens, n, fmt)
        if not res:
            raise _abi.FunasrB200Error("fa_offline_infer failed: %s" % self.lib.fa_offline_last_error().decode())
        try:
            out = []
            cnt = C.c_int32(0)
            for i in range(self.lib.fa_offline_result_count(res)):
                p = self.lib.fa_offline_result_ids(res, i, C.byref(cnt))
                out.append([int(p[k]) for k in range(cnt.value)])
            self.last_audio_seconds = float(self.lib.fa_offline_result_audio_seconds(res))
            return out
        finally:
            self.lib.fa_offline_free_result(res)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.fa_offline_uninit(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
