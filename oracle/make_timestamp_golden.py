"""Golden vectors for funasr_b200/timestamps.py: runs the REFERENCE's ts_prediction_lfr6_standard
(funasr/utils/timestamp_tools.py:37-123) on seeded CIF weights / fires and stores inputs + outputs.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (this container); the committed tests/golden/timestamps.json travels.
Usage: python oracle/make_timestamp_golden.py
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.import_reference()
from funasr.utils.timestamp_tools import ts_prediction_lfr6_standard  # noqa: E402


def main():
    rng = np.random.default_rng(20260922)
    cases = []
    for trial in range(60):
        T = int(rng.integers(8, 160))
        a = ((rng.random(T).astype(np.float32) ** 3) * 0.9).astype(np.float32)
        integ = np.float32(0)
        peaks = np.zeros(T, np.float32)
        for t in range(T):                                    # cif fires with threshold 1.0 (cif_predictor.py:853-908)
            integ = np.float32(integ + a[t])
            peaks[t] = integ
            if integ >= 1:
                integ = np.float32(integ - 1)
        ntok = int((peaks >= 1 - 1e-4).sum())
        k = max(1, ntok - 1 + (0 if trial % 3 == 0 else int(rng.integers(-1, 3))))
        chars = ["t%d" % i for i in range(k)] + (["</s>"] if trial % 5 == 1 else [])
        for up, first, second in ((1, peaks, a), (3, a, peaks)):      # the Paraformer call (model.py:674) and the BiCif-style call
            off = float((trial % 4) * 130)
            try:
                txt, res = ts_prediction_lfr6_standard(torch.tensor(first.copy()), torch.tensor(second.copy()), copy.copy(chars),
                                                       vad_offset=off, upsample_rate=up)
            except IndexError:
                txt, res = "", []
            cases.append({"first": [float(x) for x in first], "second": [float(x) for x in second], "chars": chars, "vad_offset": off,
                          "upsample_rate": up, "txt": txt, "res": res})
    out = os.path.join(ROOT, "tests", "golden", "timestamps.json")
    with open(out, "w") as f:
        json.dump(cases, f)
    print("wrote", out, len(cases), "cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
