"""Runs the dominant GEMM (FFN w_1 shape of the benchmark: M=32000, N=2048, K=512) a few times; for `ncu --set full`."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from funasr_b200 import _abi
lib = _abi.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
M, K, N = 32000, 512, 2048
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(M, K, generator=g).to(dev)
w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
b = torch.randn(N, generator=g).to(dev)
planes = torch.empty(3, N, K, dtype=torch.float16, device=dev)
st = torch.cuda.current_stream().cuda_stream
_abi.check(lib.fa_split_planes(w.data_ptr(), K, N, K, K, planes.data_ptr(), st), "split")
lin = _abi.FaLinear(w.data_ptr(), b.data_ptr(), planes.data_ptr(), N, K, K, 0)
y = torch.empty(M, N, device=dev)
ws = torch.empty(3 * M * K * 2 + 4096, dtype=torch.uint8, device=dev)
for i in range(4):
    _abi.check(lib.fa_linear(x.data_ptr(), K, M, C.byref(lin), 1, None, 0, None, 0, y.data_ptr(), N, _abi.GEMM_MODES[mode], ws.data_ptr(), ws.numel(), st), "linear")
torch.cuda.synchronize()
ref = torch.relu(x[:256] @ w.T + b)
print("probe ok, rel err %.2e" % float((y[:256] - ref).abs().max() / ref.abs().max()))
