"""In-kernel timeline of one attention CTA (debug tool, not part of the product path).

Build the trace variant first (funasr_b200/csrc/build_trace.sh -> funasr_b200/libfunasr_b200_trace.so, compiled with
-DFA_ATT_TRACE), then run this on a B200.  It calls fa_attention_tc on random q/k/v at the encoder shape and prints the
(tag, clock64) events recorded by the TMA producer, the MMA issuer and one softmax thread of a mid-grid CTA.
"""
import ctypes as C
import os
import sys

import torch

here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, "funasr_b200", "libfunasr_b200_trace.so"))
lib.fa_attention_tc_workspace_bytes.restype = C.c_size_t
lib.fa_attention_tc_workspace_bytes.argtypes = [C.c_int32] * 5
lib.fa_attention_tc.restype = C.c_int
lib.fa_attention_tc.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t,
                                C.c_void_p]
lib.fa_debug_att_trace.restype = C.c_int
lib.fa_debug_att_trace.argtypes = [C.c_void_p, C.c_void_p]

B, H, T = 64, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 500
D = H * 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, T, D, device=dev, generator=g) * 0.3
k = torch.randn(B, T, D, device=dev, generator=g)
v = torch.randn(B, T, D, device=dev, generator=g)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
ctx = torch.empty(B, T, D, device=dev)
mode = 3
ws_bytes = lib.fa_attention_tc_workspace_bytes(B, H, T, T, mode)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    rc = lib.fa_attention_tc(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, lens.data_ptr(), B, H, T, T, ctx.data_ptr(), D, mode,
                             ws.data_ptr(), ws_bytes, st)
    assert rc == 0, rc
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.fa_attention_tc(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, lens.data_ptr(), B, H, T, T, ctx.data_ptr(), D, mode, ws.data_ptr(), ws_bytes, st)
e1.record(); torch.cuda.synchronize()
print(f"fa_attention_tc (split kernels + attention) avg {e0.elapsed_time(e1) / 20 * 1000:.1f} us per call")
import numpy as np
buf = np.zeros((3, 512), dtype=np.int64)
cnt = np.zeros(3, dtype=np.int32)
assert lib.fa_debug_att_trace(buf.ctypes.data, cnt.ctypes.data) == 0
t0 = min(buf[r, 1] for r in range(3) if cnt[r] > 0)
names = ["producer", "mma", "softmax"]
for r in range(3):
    print(f"== {names[r]} ({cnt[r]} events): tag @ cycles since CTA's first event (delta)")
    prev = None
    for i in range(cnt[r]):
        tag, t = int(buf[r, 2 * i]), int(buf[r, 2 * i + 1]) - int(t0)
        print(f"  {tag:5d} @ {t:7d}" + (f"  (+{t - prev})" if prev is not None else ""))
        prev = t
ref = torch.softmax((q.view(B, T, H, 128).transpose(1, 2) @ k.view(B, T, H, 128).transpose(1, 2).transpose(-1, -2)) / 128 ** 0.5, -1) @ v.view(B, T, H, 128).transpose(1, 2)
ref = ref.transpose(1, 2).reshape(B, T, D)
print("max rel err vs torch:", float((ctx - ref).abs().max() / ref.abs().max()))
