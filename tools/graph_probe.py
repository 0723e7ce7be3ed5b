"""Measures what CUDA-graph replay of the encoder / decoder launch sequences buys over direct launches (debug tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from funasr_b200 import synth
from funasr_b200.engine import FrontendEngine, ParaformerEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda:0"
cfg = synth.ParaformerConfig()
eng = ParaformerEngine(synth.make_state_dict(cfg, 0), cfg, dev, gemm_mode="fp16x3")
fe = FrontendEngine(synth.make_cmvn(cfg, 1), dev)
base = [synth.make_wav(480000, 100 + i) for i in range(4)]
wav = torch.stack([base[i % 4].roll(977 * i) for i in range(B)]).to(dev)
wl = torch.full((B,), 480000, dtype=torch.int32, device=dev)
feats, lens = fe(wav, wl, 500)
torch.cuda.synchronize()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_direct = timeit(lambda: eng.encode(feats, lens))
# capture
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        enc = eng.encode(feats, lens)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    enc_g = eng.encode(feats, lens)
torch.cuda.synchronize()
t_graph = timeit(g.replay)
ref = eng.encode(feats, lens)
g.replay()
torch.cuda.synchronize()
print(f"encoder B={B}: direct {t_direct:.3f} ms, graph replay {t_graph:.3f} ms, max |diff| {float((ref - enc_g).abs().max()):.3e}")

acoustic, tok, _, _ = eng.predict(ref, lens)
n_max = int(tok.max())
t_dec = timeit(lambda: eng.decode(ref, lens, acoustic, tok, n_max))
s.wait_stream(torch.cuda.current_stream())
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2, stream=s):
    ids_g, _, _ = eng.decode(ref, lens, acoustic, tok, n_max)
torch.cuda.synchronize()
t_dec_g = timeit(g2.replay)
print(f"decoder n_max={n_max}: direct {t_dec:.3f} ms, graph replay {t_dec_g:.3f} ms")
