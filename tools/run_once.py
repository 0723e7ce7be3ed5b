"""Runs the hot path a few times (for ncu launch lists / captures). Usage: run_once.py --mode fp16x3 --B 16 --iters 2"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from funasr_b200 import synth
from funasr_b200.engine import FrontendEngine, ParaformerEngine
ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fp16x3"); ap.add_argument("--B", type=int, default=16); ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--layers", type=int, default=50)
a = ap.parse_args()
dev = "cuda:0"
cfg = synth.ParaformerConfig(enc_layers=a.layers, dec_layers=16 if a.layers == 50 else 2)
p = synth.make_state_dict(cfg, 0)
fe = FrontendEngine(synth.make_cmvn(cfg, 1), dev)
eng = ParaformerEngine(p, cfg, dev, gemm_mode=a.mode)
base = [synth.make_wav(480000, 100 + i) for i in range(4)]
wav = torch.stack([base[i % 4].roll(977 * i) for i in range(a.B)]).to(dev)
lens = torch.full((a.B,), 480000, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for it in range(a.iters):
    t0 = time.perf_counter()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    feats, fl = fe(wav, lens, 500)
    t1 = time.perf_counter()
    e1.record()
    out = eng.forward_feats(feats, fl)
    e2.record()
    torch.cuda.synchronize()
    print("iter %d: host fe call %.3f ms, gpu frontend %.3f ms, gpu rest %.3f ms, wall %.3f ms, tokens %d" % (
        it, (t1 - t0) * 1e3, e0.elapsed_time(e1), e1.elapsed_time(e2), (time.perf_counter() - t0) * 1e3, int(out["token_num"].sum())), flush=True)
