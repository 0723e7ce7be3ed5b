"""Writes tests/golden/knf_fbank.npz: log-mel frames produced by the reference's vendored kaldi-native-fbank, compiled from
the reference tree (oracle/knf/Makefile -> oracle/_ref/libknf_ref.so) and driven like the reference's C++ runtime drives it
(runtime/onnxruntime/src/paraformer.cpp:24-31, :298-312).  The fixture lets the tests check against that compiled reference
where the .so did not travel.  Run in the build container:  python oracle/make_knf_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import knf_ref  # noqa: E402
from funasr_b200 import synth  # noqa: E402

# (n_samples, wav seed, kind) — tests/conftest.py:KNF_CASES must match
CASES = [(16000, 31, "speechlike"), (8123, 32, "noise"), (400, 33, "noise"), (559, 34, "noise"), (48000, 35, "speechlike"), (27200, 36, "noise")]

if __name__ == "__main__":
    assert knf_ref.build(force=True), "needs /root/reference"
    out = {}
    for i, (n, seed, kind) in enumerate(CASES):
        out["logmel_%d" % i] = knf_ref.fbank(synth.make_wav(n, seed, kind).numpy())
    out["cases"] = np.array([[n, s, 0 if k == "speechlike" else 1] for n, s, k in CASES], dtype=np.int64)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "knf_fbank.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
