"""Writes tests/golden/vad_cpp_detector.npz: segments produced by the end-point detector of the reference's C++ runtime
(runtime/onnxruntime/src/e2e-vad.h, compiled from the reference tree by oracle/knf/Makefile and called like fsmn-vad.cpp:245-249)
for seeded per-frame silence posteriors.  The fixture lets tests/test_vad_host.py hold funasr_b200/vad.py to that second,
independent implementation where the compiled library is absent.  Run in the build container: python oracle/make_vad_cpp_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import knf_ref  # noqa: E402


def random_case(rng, max_seconds):
    """-> (n_samples, sil_prob [frames] fp32 on a 1/1024 grid inside (0, 1), waveform, max_end_silence_ms, speech_noise_thres)."""
    n = int(rng.uniform(0.5, max_seconds) * 16000)
    frames = 1 + (n - 400) // 160
    sp = np.empty(frames, np.float32)
    t, speech = 0, bool(rng.integers(0, 2))
    while t < frames:
        length = int(rng.integers(3, 400))
        level = rng.uniform(0.0, 0.35) if speech else rng.uniform(0.65, 1.0)
        sp[t: t + length] = np.clip(np.round((level + rng.normal(0, 0.15, size=min(length, frames - t))) * 1024) / 1024, 1 / 1024, 1023 / 1024)
        t += length
        speech = not speech
    return n, sp, flat_wave(n), int(rng.choice([400, 800, 1200])), float(rng.choice([0.6, 0.8]))


def flat_wave(n):
    """Deterministic +-0.05 square wave: every frame has the same energy, far above the detector's -100 dB floors."""
    w = np.full(n, 0.05, np.float32)
    w[1::2] = -0.05
    return w


if __name__ == "__main__":
    assert knf_ref.build(force=True), "needs /root/reference"
    rng = np.random.default_rng(20260923)
    out = {}
    meta = []
    for i in range(16):
        n, sp, wav, mes, thr = random_case(rng, 40.0)
        seg = knf_ref.vad_segments(sp, wav, mes, 60000, thr)
        out["sil_prob_%d" % i] = (sp * 1024).astype(np.int16)
        out["segments_%d" % i] = np.array(seg, dtype=np.int64).reshape(-1, 2)
        meta.append([n, mes, int(round(thr * 10))])
    out["meta"] = np.array(meta, dtype=np.int64)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "vad_cpp_detector.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", sum(len(out["segments_%d" % i]) for i in range(16)), "segments")
