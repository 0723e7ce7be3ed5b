"""Where does the arithmetic noise of the tensor-core path come from?  (writes gpurun_out/noise_probe.json)
  1. one GEMM (M=4096, N=2048, K=512 and K=2048) against a float64 reference: max / rms relative error and the SIGNED bias
     (mean of (y - ref) * sign(ref) / mean |ref|: a round-toward-zero accumulator shows up as a negative bias) for fp16 / x3 / x6;
  2. the 50-layer encoder and the log-probs of 2 x 30 s in fp16x3 and fp16x6 against the fp32 SIMT path on the same GPU."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from funasr_b200 import _abi, synth  # noqa: E402
from funasr_b200.engine import FrontendEngine, ParaformerEngine  # noqa: E402

dev = "cuda:0"
lib = _abi.load()
out = {"gemm": [], "model": {}}
g = torch.Generator().manual_seed(0)
for K in (512, 2048):
    M, N = 4096, 2048 if K == 512 else 512
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    ref = (x.double() @ w.double().t())
    xd, wd = x.to(dev), w.to(dev)
    planes = torch.empty(3, N, K, dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _abi.check(lib.fa_split_planes(wd.data_ptr(), K, N, K, K, planes.data_ptr(), st), "split")
    lin = _abi.FaLinear(wd.data_ptr(), None, planes.data_ptr(), N, K, K, 0)
    ws = torch.empty(3 * M * K * 2 + 4096, dtype=torch.uint8, device=dev)
    for mode in ("fp32", "fp16", "fp16x3", "fp16x6"):
        y = torch.empty(M, N, device=dev)
        _abi.check(lib.fa_linear(xd.data_ptr(), K, M, C.byref(lin), 0, None, 0, None, 0, y.data_ptr(), N, _abi.GEMM_MODES[mode], ws.data_ptr(), ws.numel(), st), mode)
        d = y.double().cpu() - ref
        scale = ref.abs().mean()
        row = {"K": K, "mode": mode, "max_rel": float(d.abs().max() / ref.abs().max()), "rms_rel": float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
               "signed_bias": float((d * ref.sign()).mean() / scale)}
        out["gemm"].append(row)
        print(row, flush=True)
cfg = synth.PARAFORMER_LARGE
p = synth.make_state_dict(cfg, 0)
fe = FrontendEngine(synth.make_cmvn(cfg, 1), dev)
wavs = [synth.make_wav(480000, 1000, "speechlike"), synth.make_wav(480000, 1001, "speechlike")]
feats, fl = fe(torch.stack(wavs).to(dev), torch.full((2,), 480000, dtype=torch.int32, device=dev), 500)
res = {}
for mode in ("fp32", "fp16x3", "fp16x6"):
    eng = ParaformerEngine(p, cfg, dev, gemm_mode=mode)
    o = eng.forward_feats(feats, fl, want_taps=True)
    res[mode] = (o["enc"].double().cpu(), o["logp"].double().cpu(), o["ids"])
    del eng
    torch.cuda.empty_cache()
for mode in ("fp16x3", "fp16x6"):
    e = (res[mode][0] - res["fp32"][0]).abs().max() / res["fp32"][0].abs().max()
    n = min(res[mode][1].shape[1], res["fp32"][1].shape[1])
    lp = (res[mode][1][:, :n] - res["fp32"][1][:, :n]).abs().max() / res["fp32"][1][:, :n].abs().max()
    out["model"][mode] = {"enc_rel_err_vs_fp32": float(e), "logp_rel_err_vs_fp32": float(lp), "ids_equal": res[mode][2] == res["fp32"][2]}
    print(mode, out["model"][mode], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("NOISE_PROBE_OUT", "noise_probe.json")), "w"), indent=1)
