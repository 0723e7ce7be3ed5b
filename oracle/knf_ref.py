"""TEST INFRASTRUCTURE: loader for oracle/_ref/libknf_ref.so — the reference's vendored kaldi-native-fbank compiled from
the reference tree (oracle/knf/Makefile), configured and fed as the reference's C++ runtime does
(runtime/onnxruntime/src/paraformer.cpp:24-31, :298-312).  Second pin of the Fbank arithmetic, independent of torchaudio.
Only tests/, __graft_entry__.build()/smoke() and bench.py's CPU leg may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libknf_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(force: bool = False) -> bool:
    """Compile from the reference tree when it is present (this container); the GPU box uses the prebuilt file."""
    if os.path.exists(SO) and not force:
        return True
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "runtime", "onnxruntime", "third_party", "kaldi-native-fbank")):
        return False
    r = subprocess.run(["make", "-C", os.path.join(HERE, "knf"), "REF=" + REFERENCE_ROOT] + (["-B"] if force else []),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/knf build failed:\n" + r.stdout[-2000:])
    return os.path.exists(SO)


_lib = None


def available() -> bool:
    return os.path.exists(SO)


def fbank(wav: np.ndarray, fs: float = 16000.0, n_mels: int = 80, frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0) -> np.ndarray:
    """wav: float32 in [-1, 1] (the driver scales by 32768 like Paraformer::FbankKaldi) -> log-mel [frames, n_mels]."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.knf_ref_fbank.restype = C.c_int
        _lib.knf_ref_fbank.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int]
    a = np.ascontiguousarray(wav, dtype=np.float32)
    nf = _lib.knf_ref_fbank(a.ctypes.data, a.size, fs, n_mels, frame_length_ms, frame_shift_ms, None, 0)
    out = np.empty((max(nf, 0), n_mels), np.float32)
    if nf > 0:
        _lib.knf_ref_fbank(a.ctypes.data, a.size, fs, n_mels, frame_length_ms, frame_shift_ms, out.ctypes.data, nf)
    return out
