"""Times the pair-tile tcgen05 GEMM (gemm_tc2_kernel) alone on the shapes the Paraformer step launches, L2 flushed between
launches, CUDA events on the launching stream.  Prints one JSON object (also written to gpurun_out/gemm_shapes[_TAG].json).

    python tools/gemm_shapes.py [--tag NAME]          FA_GEMM_TAIL=0 python tools/gemm_shapes.py --tag whole   (A/B of the tail slices)

Every output is also checked against torch fp64 products of the SAME fp16 planes (hi*hi + hi*lo + lo*hi is exactly what the
kernel sums), so a slice-indexing bug in the tail schedule shows up as an O(1) error rather than as noise."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from funasr_b200 import _abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="")
ap.add_argument("--mode", default="fp16x3")
ap.add_argument("--only", default="", help="substring filter on the shape names")
args = ap.parse_args()
dev = "cuda:0"
lib = _abi.load()
gm = _abi.GEMM_MODES[args.mode]
npl = {"fp16": 1, "fp16x3": 2}[args.mode]
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)

# (name, M, N, K, kind)   kind: "f32" = fp32 rows + one residual (out-projection, FFN w_2), "planes" = ReLU fp16 planes (FFN w_1)
SHAPES = [("enc out-proj", 32000, 512, 512, "f32"), ("enc w_1", 32000, 2048, 512, "planes"), ("enc w_2", 32000, 512, 2048, "f32"),
          ("enc qkv-like", 32000, 1536, 512, "planes"),
          ("dec q/out", 9984, 512, 512, "f32"), ("dec w_1", 9984, 2048, 512, "planes"), ("dec w_2", 9984, 512, 2048, "f32"),
          ("dec kv", 32000, 1024, 512, "planes"), ("dec ragged rows", 10007, 512, 512, "f32"),
          ("B=1 out-proj", 500, 512, 512, "f32"), ("B=1 w_1", 500, 2048, 512, "planes"),
          # diagnostics: the same out-projection without its residual stream / with plane output (what bounds the epilogue?)
          ("diag out-proj no residual", 32000, 512, 512, "f32-nores"), ("diag out-proj planes", 32000, 512, 512, "planes"),
          # ragged N on pair tiles: the vocabulary projections (Paraformer 8404, SenseVoice 25055) and a ragged M x ragged N corner
          ("vocab 8404", 9984, 8404, 512, "f32-nores"), ("vocab 25055", 8192, 25055, 512, "f32-nores"), ("ragged both", 1000, 1300, 512, "f32-nores")]
SHAPES = [s_ for s_ in SHAPES if args.only in s_[0]]
out = {"mode": args.mode, "tail": os.environ.get("FA_GEMM_TAIL", "1"), "shapes": []}
for name, M, N, K, kind in SHAPES:
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    wp = torch.empty(3, N, K, dtype=torch.float16, device=dev)
    _abi.check(lib.fa_split_planes(w.data_ptr(), K, N, K, K, wp.data_ptr(), st), "split w")
    lin = _abi.FaLinear(w.data_ptr(), b.data_ptr(), wp.data_ptr(), N, K, K, 0)
    xp = torch.empty(npl, M, K, dtype=torch.float16, device=dev)
    _abi.check(lib.fa_split_rows(x.data_ptr(), K, M, K, K, npl, xp.data_ptr(), st), "split x")
    res = torch.randn(M, N, generator=g).to(dev) if kind == "f32" else None
    y = torch.empty(M, N, device=dev) if kind.startswith("f32") else None
    yp = torch.empty(npl, M, N, dtype=torch.float16, device=dev) if kind == "planes" else None

    def launch():
        if kind.startswith("f32"):
            _abi.check(lib.fa_linear_planes(xp.data_ptr(), M, C.byref(lin), 0, res.data_ptr() if res is not None else None, N, None, 0, y.data_ptr(), N, gm, st),
                       "fa_linear_planes")
        else:
            _abi.check(lib.fa_linear_planes_to_planes(xp.data_ptr(), M, C.byref(lin), 1, yp.data_ptr(), N, gm, st), "fa_linear_planes_to_planes")

    times = []
    for i in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1) * 1e3)
    # correctness against fp64 on the SAME planes (hi*hi + hi*lo + lo*hi is what the kernel sums)
    xs, ws = xp.double(), wp.double()
    if npl == 2:
        ref = xs[0] @ ws[0].t() + xs[0] @ ws[1].t() + xs[1] @ ws[0].t()
    else:
        ref = xs[0] @ ws[0].t()
    ref = ref + b.double()
    if kind.startswith("f32"):
        got = y.double(); ref = ref + (res.double() if res is not None else 0)
    else:
        got = yp.double().sum(0); ref = ref.clamp_min(0)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    tiles = -(-M // 256) * (-(-N // 256))
    row = {"name": name, "M": M, "N": N, "K": K, "kind": kind, "us": sum(times) / len(times), "us_min": min(times), "tiles": tiles, "rounds": tiles / 74.0,
           "tflops_issue": 2.0 * M * N * K * (3 if npl == 2 else 1) / (sum(times) / len(times)) / 1e6, "max_rel_err_vs_fp64": err}
    print(row, flush=True)
    assert err < 2e-5, (name, err)
    out["shapes"].append(row)
fn = "gemm_shapes%s.json" % (("_" + args.tag) if args.tag else "")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", fn), "w"), indent=1)
