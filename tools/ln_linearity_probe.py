"""CPU numerics check for DESIGN.md §7b item 3 (not a product path): the decoder FFN's LayerNorm folded into w_2 by linearity,
    w_2 . LN(h) = rstd * (h (W_2 gamma)^T - mu * sum_k(W_2 gamma)) + W_2 beta,
so that w_1 could emit ReLU planes + row statistics and the 2048-wide LayerNorm pass disappears.  The subtraction moves behind the
GEMM: this script measures what that costs in fp32 against an fp64 reference, next to the direct fp32 route, on the synthetic
decoder weights (funasr_b200/synth.py).  No GPU."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from funasr_b200 import synth  # noqa: E402

torch.manual_seed(0)
p = synth.make_state_dict(synth.PARAFORMER_LARGE, 0)
rep = {}
for layer in (0, 3, 15):
    pre = "decoder.decoders.%d.feed_forward." % layer
    W1, b1 = p[pre + "w_1.weight"].double(), p[pre + "w_1.bias"].double()
    g, be, W2 = p[pre + "norm.weight"].double(), p[pre + "norm.bias"].double(), p[pre + "w_2.weight"].double()
    x = torch.nn.functional.layer_norm(torch.randn(4096, 512, dtype=torch.float64), (512,))
    h = torch.relu(x @ W1.T + b1)
    ref = torch.nn.functional.layer_norm(h, (2048,), g, be, 1e-12) @ W2.T
    h32 = h.float()
    Wg = (W2 * g).float()
    sw, wb = Wg.double().sum(1).float(), (W2 @ be).float()
    mu = h32.mean(1, keepdim=True)
    rstd = torch.rsqrt((h32 - mu).pow(2).mean(1, keepdim=True) + 1e-12)
    y = rstd * (h32 @ Wg.T - mu * sw) + wb
    y_dir = torch.nn.functional.layer_norm(h32, (2048,), g.float(), be.float(), 1e-12) @ W2.float().T
    den = float(ref.abs().max())
    rep["decoder layer %d" % layer] = {"linearity_route_rel_err": float((y.double() - ref).abs().max()) / den,
                                       "direct_fp32_rel_err": float((y_dir.double() - ref).abs().max()) / den,
                                       "cancelled_term_over_result": float((mu * sw).abs().mean() / ref.abs().mean())}
print(json.dumps(rep, indent=1))
