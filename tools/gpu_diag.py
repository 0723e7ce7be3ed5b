"""One-shot GPU diagnostics: per-stage parity numbers against the golden vectors and stage timings.
Writes gpurun_out/diag_<tag>.json.  Usage: python tools/gpu_diag.py [tag] [mode] [--time B]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import GOLDEN_CASES, load_case, rel_err, state_dict_for  # noqa: E402
from funasr_b200 import synth  # noqa: E402
from funasr_b200.engine import FrontendEngine, ParaformerEngine, num_lfr_frames  # noqa: E402

DEV = "cuda:0"


def parity(mode):
    rep = {}
    for name in GOLDEN_CASES:
        cfg, wseed, wavs, cmvn, g = load_case(name)
        fe = FrontendEngine(cmvn, DEV)
        eng = ParaformerEngine(state_dict_for(cfg, wseed), cfg, DEV, gemm_mode=mode)
        lens = [w.numel() for w in wavs]
        pad = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True).to(DEV)
        feats, fl = fe(pad, torch.tensor(lens, dtype=torch.int32, device=DEV), max(num_lfr_frames(n) for n in lens))
        o = eng.forward_feats(feats, fl, want_taps=True)
        torch.cuda.synchronize()
        big = cfg.enc_layers > 10
        sub = lambda t, s: (t[:, ::s] if big else t)
        n = int(g["token_num"].max())
        r = {"feat_lens_ok": fl.cpu().tolist() == g["feat_lens"].tolist(),
             "feats_maxabs": float(np.abs(sub(feats.cpu(), 7).numpy() - g["feats"]).max()),
             "enc_rel": rel_err(sub(o["enc"].cpu(), 7).numpy(), g["enc"]),
             "alphas_maxabs": float(np.abs(o["alphas"].cpu().numpy() - g["alphas"]).max()),
             "token_num": o["token_num"].tolist(), "token_num_ref": g["token_num"].tolist()}
        if o["token_num"].tolist() == g["token_num"].tolist() and "logp" in o:
            r["acoustic_rel"] = rel_err(sub(o["acoustic"][:, :n].cpu(), 5).numpy(), g["acoustic"])
            lp = o["logp"][:, g["logp_rows"].tolist(), :].cpu().numpy()
            r["logp_rel"] = rel_err(lp, g["logp_sel"])
            r["logp_maxabs"] = float(np.abs(lp - g["logp_sel"]).max())
            valid = np.arange(n)[None, :] < g["token_num"][:, None]
            am = o["argmax"].cpu().numpy()
            r["argmax_mismatch"] = int((am[valid] != g["argmax"][valid]).sum())
            r["n_tokens"] = int(valid.sum())
            r["min_margin_ref"] = float(g["margin"][valid].min())
            r["ids_equal"] = [t for x in o["ids"] for t in x] == g["ids_flat"].tolist()
        rep[name] = r
        print(name, json.dumps(r), flush=True)
        del eng
        torch.cuda.empty_cache()
    return rep


def timing(mode, B):
    cfg = synth.PARAFORMER_LARGE
    p = state_dict_for(cfg, 0)
    cmvn = synth.make_cmvn(cfg, 1)
    fe = FrontendEngine(cmvn, DEV)
    eng = ParaformerEngine(p, cfg, DEV, gemm_mode=mode)
    base = [synth.make_wav(480000, 100 + i) for i in range(4)]
    wav = torch.stack([base[i % 4].roll(977 * i) for i in range(B)]).to(DEV)
    lens = torch.full((B,), 480000, dtype=torch.int32, device=DEV)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    res = {}
    for it in range(3):
        e = [ev() for _ in range(6)]
        e[0].record()
        feats, fl = fe(wav, lens, 500)
        e[1].record()
        enc = eng.encode(feats, fl)
        e[2].record()
        ac, tok, al, pk = eng.predict(enc, fl)
        e[3].record()
        tok_h = tok.cpu()
        n_max = int(tok_h.max())
        e[4].record()
        ids, best, _ = eng.decode(enc, fl, ac, tok, n_max)
        fids, flens = eng.greedy_filter(ids, tok)
        e[5].record()
        torch.cuda.synchronize()
        res = {"frontend_ms": e[0].elapsed_time(e[1]), "encoder_ms": e[1].elapsed_time(e[2]), "predictor_ms": e[2].elapsed_time(e[3]),
               "decoder_ms": e[4].elapsed_time(e[5]), "total_ms": e[0].elapsed_time(e[5]), "n_max": n_max,
               "tokens_mean": float(tok_h.float().mean()), "B": B, "mode": mode}
        res["rtfx"] = B * 30.0 / (res["total_ms"] / 1000)
        print("timing", json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    out = {"parity": parity(mode)}
    if "--time" in sys.argv:
        out["timing"] = timing(mode, int(sys.argv[sys.argv.index("--time") + 1]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_%s_%s.json" % (tag, mode)), "w"), indent=1)
