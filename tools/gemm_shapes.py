"""Times fa_linear_planes (the tcgen05 GEMM with pre-split operand planes) at the layer shapes of the benchmark, with and
without residual streams, against the 3-pass tensor-issue floor.  Debug / measurement tool."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from funasr_b200 import _abi

lib = _abi.load()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
gm = _abi.GEMM_MODES["fp16x3"]
M = 32000
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peak = 1708.3e12


def run(name, N, K, n_res, relu=0, iters=8):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    wp = torch.empty(3, N, K, dtype=torch.float16, device=dev)
    _abi.check(lib.fa_split_planes(w.data_ptr(), K, N, K, K, wp.data_ptr(), st), "split")
    lin = _abi.FaLinear(w.data_ptr(), b.data_ptr(), wp.data_ptr(), N, K, K, 0)
    xp = torch.empty(2, M, K, dtype=torch.float16, device=dev)
    _abi.check(lib.fa_split_rows(x.data_ptr(), K, M, K, K, 2, xp.data_ptr(), st), "split_rows")
    y = torch.empty(M, N, device=dev)
    r1 = torch.randn(M, N, device=dev) if n_res >= 1 else None
    r2 = torch.randn(M, N, device=dev) if n_res >= 2 else None
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.check(lib.fa_linear_planes(xp.data_ptr(), M, C.byref(lin), relu, r1.data_ptr() if r1 is not None else None, N,
                                        r2.data_ptr() if r2 is not None else None, N, y.data_ptr(), N, gm, st), "fa_linear_planes")
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    us = sum(ts) / len(ts)
    floor = 3 * 2.0 * M * N * K / peak * 1e6
    byts = (2 * M * K * 2 + 2 * N * K * 2 + M * N * 4 * (1 + n_res))
    print(f"{name:28s} N={N:5d} K={K:5d} res={n_res}: {us:7.1f} us  (3-pass tensor floor {floor:6.1f} us, HBM floor {byts / 6.5648e12 * 1e6:6.1f} us)")


run("out-proj, 2 residuals", 512, 512, 2)
run("out-proj, 1 residual", 512, 512, 1)
run("out-proj, no residual", 512, 512, 0)
run("w_2, 1 residual", 512, 2048, 1)
run("w_2, no residual", 512, 2048, 0)
run("w_1 shape (fp32 out)", 2048, 512, 0, relu=1)
run("N=1024 K=512", 1024, 512, 0)
run("N=1536 K=512 (QKV shape, fp32 out)", 1536, 512, 0)
