"""Stage-by-stage comparison of the CT-Transformer GPU path with the oracle (run by hand on a GPU box: python tools/punc_diag.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import paraformer_oracle as O  # noqa: E402
from funasr_b200 import _abi, synth  # noqa: E402
from funasr_b200.punc import PuncEngine  # noqa: E402

p = synth.make_punc_state_dict(0)
eng = PuncEngine(p, torch.device("cuda:0"), synth.PUNC_HEADS)
ids = np.arange(5, 5 + 37, dtype=np.int32)
T = len(ids)
x_ref = torch.nn.functional.embedding(torch.from_numpy(ids.astype(np.int64))[None], p["embed.weight"])
h_ref, _ = O.encoder(x_ref, torch.tensor([T]), p, synth.PUNC_LAYERS, heads=synth.PUNC_HEADS)
lg_ref = torch.nn.functional.linear(h_ref, p["decoder.weight"], p["decoder.bias"])[0]
dev = eng.device
idt = torch.from_numpy(ids).to(dev)
x = torch.empty((1, T, eng.d_in), dtype=torch.float32, device=dev)
st = eng._stream()
_abi.check(eng.lib.fa_embedding(idt.data_ptr(), eng.embed.data_ptr(), eng.d_in, int(eng.embed.shape[0]), T, x.data_ptr(), st), "emb")
print("embedding max err", (x.cpu() - x_ref).abs().max().item())
lens = torch.tensor([T], dtype=torch.int32, device=dev)
h = eng._encode(eng.enc, x, lens, eng.d_model)
torch.cuda.synchronize()
print("encoder rel err", ((h.cpu() - h_ref).abs().max() / h_ref.abs().max()).item())
for nl in range(1, synth.PUNC_LAYERS + 1):
    pass
got = eng.punc_ids(ids)
print("ids", got.tolist())
print("ref", lg_ref.argmax(-1).tolist())
