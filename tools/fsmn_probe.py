"""A/B of the two FSMN memory-block kernels at the benchmark's encoder shape (B = 64, T = 500, 512 channels, v = the fp32 V columns
of the QKV buffer, fused residual): the SIMT strip kernel (fa_fsmn_simt) against the TMA-staged warp-specialised one (fa_fsmn_tma).
CUDA events on the launching stream, L2 flushed before every launch (and, second column, back to back without a flush: the state
the kernel sees in the step, where QKV / the residual stream were just written).  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from funasr_b200 import _abi  # noqa: E402

lib = _abi.load()
dev = "cuda:0"
B, T, Cn, LD, K = 64, 500, 512, 1536, 11
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, LD, generator=g).to(dev)
w = (torch.randn(Cn, 1, K, generator=g) * 0.2).to(dev)
res = torch.randn(B, T, Cn, generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
out = {"simt": torch.empty(B, T, Cn, device=dev), "tma": torch.empty(B, T, Cn, device=dev)}
fn = {"simt": lib.fa_fsmn_simt, "tma": lib.fa_fsmn_tma}
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
algo_bytes = 3 * B * T * Cn * 4


def run(name):
    _abi.check(fn[name](qkv.data_ptr() + 1024 * 4, LD, lens.data_ptr(), B, T, Cn, w.data_ptr(), K, res.data_ptr(), Cn, out[name].data_ptr(), Cn, st), name)


def timed(name, do_flush, n=20):
    ts = []
    for i in range(n + 3):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(name)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rep = {"shape": {"B": B, "T": T, "channels": Cn, "ldv": LD, "k": K, "residual": True}, "algorithmic_bytes": algo_bytes}
for name in ("simt", "tma"):
    cold, warm = timed(name, True), timed(name, False)
    rep[name] = {"us_l2_flushed": cold, "us_back_to_back": warm, "gbs_l2_flushed": algo_bytes / cold / 1e3, "gbs_back_to_back": algo_bytes / warm / 1e3}
rep["bit_identical"] = bool(torch.equal(out["simt"], out["tma"]))
print(json.dumps(rep))
