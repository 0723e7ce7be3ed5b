"""TEST INFRASTRUCTURE: loader for the compiled-reference checkers under oracle/_ref/ (oracle/knf/Makefile builds them from the
reference tree where it lies):
  * libknf_ref.so — the reference's vendored kaldi-native-fbank, configured and fed as the reference's C++ runtime does
    (runtime/onnxruntime/src/paraformer.cpp:24-31, :298-312): second pin of the Fbank arithmetic, independent of torchaudio;
  * libvad_ref.so — the end-point detector of the C++ runtime (runtime/onnxruntime/src/e2e-vad.h, called like fsmn-vad.cpp:245-249):
    an implementation of the VAD state machine independent of the Python model that funasr_b200/vad.py restates (it has no dynamic
    end-silence schedule: compared with a fixed max_end_silence_time).
Only tests/, __graft_entry__.build()/smoke() and bench.py's CPU leg may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libknf_ref.so")
VAD_SO = os.path.join(HERE, "_ref", "libvad_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(force: bool = False) -> bool:
    """Compile from the reference tree when it is present (this container); the GPU box uses the prebuilt file."""
    if os.path.exists(SO) and os.path.exists(VAD_SO) and not force:
        return True
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "runtime", "onnxruntime", "third_party", "kaldi-native-fbank")):
        return False
    r = subprocess.run(["make", "-C", os.path.join(HERE, "knf"), "REF=" + REFERENCE_ROOT] + (["-B"] if force else []),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/knf build failed:\n" + r.stdout[-2000:])
    return os.path.exists(SO) and os.path.exists(VAD_SO)


_lib = None
_vad_lib = None


def vad_segments(sil_prob, wav, max_end_silence_ms: int = 800, max_single_segment_ms: int = 60000, speech_noise_thres: float = 0.6,
                 sample_rate: int = 16000):
    """Per-frame silence posteriors + the waveform -> [[start_ms, end_ms], ...] from the C++ runtime's E2EVadModel (offline, is_final)."""
    global _vad_lib
    if _vad_lib is None:
        _vad_lib = C.CDLL(VAD_SO)
        _vad_lib.vad_ref_segments.restype = C.c_int
        _vad_lib.vad_ref_segments.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
    sp = np.ascontiguousarray(sil_prob, dtype=np.float32)
    w = np.ascontiguousarray(wav, dtype=np.float32)
    cap = 64
    while True:
        out = np.zeros((cap, 2), np.int32)
        n = _vad_lib.vad_ref_segments(sp.ctypes.data, sp.size, w.ctypes.data, w.size, int(max_end_silence_ms), int(max_single_segment_ms),
                                      float(speech_noise_thres), int(sample_rate), out.ctypes.data, cap)
        if n <= cap:
            return out[:n].tolist()
        cap = n


def available() -> bool:
    return os.path.exists(SO) and os.path.exists(VAD_SO)


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
