"""Per-kernel roofline table of one 64 x 30 s forward from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none
--csv ... python tools/run_once.py --B 64 --iters 1`).  Every launch is mapped to its operation by its position in the (fixed) launch
sequence of the encoder / predictor / decoder; the shapes follow from the configuration and from the decoder LayerNorm's grid
(rows = B * n_max).  Work is ALGORITHMIC (2MNK flops of the fp32-equivalent product, bytes each element is read / written once);
peaks are the measured ones in MEASURED_PEAKS.json (burst: each kernel runs alone under ncu).  Usage:
    python tools/step_roofline.py profiles/r2_launches_fsmn_tma.csv > profiles/r2_step_roofline.md"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, T, D, F, H, V, K560 = 64, 500, 512, 2048, 4, 8404, 560


def load(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    idx = {h: i for i, h in enumerate(rows[0])}
    out = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("fa::", "")
        us = float(r[idx["Metric Value"]]) / (1e3 if r[idx["Metric Unit"]] == "ns" else 1.0)
        grid = int(re.match(r"\((\d+)", r[idx["Grid Size"]]).group(1))
        out.append((name, us, grid))
    start = max(i for i, s in enumerate(out) if s[0].startswith("split_planes")) + 1     # after the one-time weight split
    while not out[start][0].startswith("fbank_tab"):
        start += 1
    return out[start:]


def main():
    seq = load(sys.argv[1])
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak, hbm_peak = pk.get("bf16_tflops", 1708.3), pk.get("hbm_gbs", 6564.8)
    M = B * T
    cif = next(i for i, s in enumerate(seq) if s[0].startswith("cif_pad"))
    dec_ln = next(s for s in seq[cif:] if s[0].startswith("layernorm"))
    Md = dec_ln[2] * 8                                   # LayerNorm: 8 rows (warps) per CTA
    n_max = Md // B
    ops = collections.OrderedDict()

    def add(key, us, flops=0.0, bytes_=0.0, bound="tensor"):
        o = ops.setdefault(key, {"n": 0, "us": 0.0, "flops": 0.0, "bytes": 0.0, "bound": bound})
        o["n"] += 1; o["us"] += us; o["flops"] += flops; o["bytes"] += bytes_

    # ---- encoder: fbank, then per layer LN, QKV, FSMN, attention, out-proj, LN, w_1, w_2; final LN
    i = 0
    add("fbank + LFR + CMVN (fbank_tab_kernel)", seq[i][1], bytes_=B * (480000 * 4 + T * 560 * 4), bound="hbm (issue bound, DESIGN §6)"); i += 1
    layer = 0
    while i < cif - 1:
        kin = K560 if layer == 0 else D
        add("enc LayerNorm -> 2 fp16 planes", seq[i][1], bytes_=M * kin * 4 + 2 * M * ((kin + 63) // 64 * 64) * 2, bound="hbm"); i += 1
        assert seq[i][0].startswith("gemm_tc2_kernel<3, 2, 2>"), seq[i]
        add("enc QKV GEMM (-> q/k planes, v^T planes, fp32 v)", seq[i][1], flops=2.0 * M * 3 * D * kin); i += 1
        assert seq[i][0].startswith("fsmn"), seq[i]
        add("enc FSMN memory (+ residual)", seq[i][1], bytes_=(3 if layer else 2) * M * D * 4, bound="hbm"); i += 1
        assert seq[i][0].startswith("attention"), seq[i]
        add("enc self-attention (QK^T + softmax + PV)", seq[i][1], flops=4.0 * B * T * T * D); i += 1
        add("enc out-projection GEMM (+ residual, fp32)", seq[i][1], flops=2.0 * M * D * D); i += 1
        add("enc LayerNorm -> 2 fp16 planes", seq[i][1], bytes_=M * D * 4 + 2 * M * D * 2, bound="hbm"); i += 1
        assert seq[i][0].startswith("gemm_tc2_kernel<3, 2, 1>"), seq[i]
        add("enc FFN w_1 GEMM (ReLU -> 2 fp16 planes)", seq[i][1], flops=2.0 * M * F * D); i += 1
        add("enc FFN w_2 GEMM (+ residual, fp32)", seq[i][1], flops=2.0 * M * D * F); i += 1
        layer += 1
    add("enc after_norm (fp32)", seq[i][1], bytes_=2 * M * D * 4, bound="hbm"); i += 1
    # ---- predictor
    add("CIF pad planes", seq[i][1], bytes_=M * D * 4 + 2 * M * D * 2, bound="hbm"); i += 1
    add("CIF conv k=3 as one GEMM (overlapping TMA view)", seq[i][1], flops=2.0 * M * D * 3 * D); i += 1
    add("CIF alpha head", seq[i][1], bytes_=M * D * 4, bound="hbm"); i += 1
    add("CIF scan + fire + segment sums", seq[i][1], bytes_=M * D * 4 + B * n_max * D * 4, bound="hbm (latency chain)"); i += 1
    # ---- decoder
    add("dec enc -> planes (split_rows)", seq[i][1], bytes_=M * D * 4 + 2 * M * D * 2, bound="hbm"); i += 1
    dl = 0
    while seq[i + 1][0].startswith("gemm") and seq[i + 2][0].startswith("layernorm_kernel<16"):
        full = seq[i + 5][0].startswith("fsmn")
        add("dec LayerNorm 512 -> planes / fp32", seq[i][1], bytes_=Md * D * 8, bound="hbm"); i += 1
        add("dec FFN w_1 GEMM (ReLU, fp32)", seq[i][1], flops=2.0 * Md * F * D); i += 1
        add("dec FFN LayerNorm 2048 -> planes", seq[i][1], bytes_=Md * F * 4 + 2 * Md * F * 2, bound="hbm"); i += 1
        add("dec FFN w_2 GEMM (+ residual)", seq[i][1], flops=2.0 * Md * D * F); i += 1
        if not full:
            break
        add("dec LayerNorm 512 -> planes / fp32", seq[i][1], bytes_=Md * D * 8, bound="hbm"); i += 1
        add("dec FSMN memory (+ residual)", seq[i][1], bytes_=3 * Md * D * 4, bound="hbm"); i += 1
        add("dec LayerNorm 512 -> planes / fp32", seq[i][1], bytes_=Md * D * 8, bound="hbm"); i += 1
        add("dec cross-attention q GEMM", seq[i][1], flops=2.0 * Md * D * D); i += 1
        add("dec cross-attention k/v GEMM (encoder rows)", seq[i][1], flops=2.0 * M * 2 * D * D); i += 1
        add("dec cross-attention", seq[i][1], flops=4.0 * B * n_max * T * D); i += 1
        add("dec cross-attention out GEMM (+ residual)", seq[i][1], flops=2.0 * Md * D * D); i += 1
        dl += 1
    rest = seq[i:]
    for name, us, grid in rest:
        if name.startswith("gemm"):
            add("vocabulary projection GEMM (N = 8404)", us, flops=2.0 * Md * V * D)
        elif name.startswith("argmax"):
            add("log-softmax arg-max over the vocabulary", us, bytes_=Md * V * 4, bound="hbm")
        elif name.startswith("layernorm"):
            add("dec LayerNorm 512 -> planes / fp32", us, bytes_=Md * D * 8, bound="hbm")
        else:
            add("other (" + name.split("<")[0] + ")", us, bound="-")
    tot = sum(o["us"] for o in ops.values())
    print("# Step roofline of one 64 x 30 s forward (fp16x3 parity mode), from `%s`\n" % os.path.basename(sys.argv[1]))
    print("%d launches, %.2f ms of kernel time (each kernel alone under ncu: cold L2, no power cap — the timed step is 42-43 ms), decoder rows "
          "B x n_max = 64 x %d, %d encoder + %d decoder layers.  Peaks: %.1f TFLOP/s 16-bit dense (burst), %.1f GB/s HBM (MEASURED_PEAKS.json).  "
          "`frac` = algorithmic work / time / peak; tensor-bound kernels of the x3 split issue 3 MMAs per product, so their tensor pipe runs at "
          "`3 x frac` (attention: 7/6 x 3 = 3.5 — pass A adds a seventh MMA unit per key chunk to the six of the exact products).\n" % (
              len(seq), tot / 1e3, n_max, layer, dl, tf_peak, hbm_peak))
    print("| operation | launches | avg µs | total ms | share | bound | algorithmic work / launch | achieved | frac of peak | x3 issue frac |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for k, o in ops.items():
        avg = o["us"] / o["n"]
        if o["flops"]:
            ach = o["flops"] / o["us"] / 1e6
            mult = 3 * 7 / 6 if "attention" in k and "GEMM" not in k else 3
            print("| %s | %d | %.1f | %.2f | %.1f %% | tensor | %.1f GFLOP | %.0f TFLOP/s | %.3f | %.2f |" % (
                k, o["n"], avg, o["us"] / 1e3, 100 * o["us"] / tot, o["flops"] / o["n"] / 1e9, ach, ach / tf_peak, mult * ach / tf_peak))
        elif o["bytes"]:
            ach = o["bytes"] / o["us"] / 1e3
            print("| %s | %d | %.1f | %.2f | %.1f %% | %s | %.1f MB | %.0f GB/s | %.3f | |" % (
                k, o["n"], avg, o["us"] / 1e3, 100 * o["us"] / tot, o["bound"], o["bytes"] / o["n"] / 1e6, ach, ach / hbm_peak))
        else:
            print("| %s | %d | %.1f | %.2f | %.1f %% | | | | | |" % (k, o["n"], avg, o["us"] / 1e3, 100 * o["us"] / tot))
    fl = sum(o["flops"] for o in ops.values())
    tus = sum(o["us"] for o in ops.values() if o["flops"])
    print("\nTensor-bound kernels together: %.2f TFLOP algorithmic in %.2f ms = %.0f TFLOP/s = %.3f of the burst peak (x3 issue: %.2f); "
          "HBM-bound kernels together: %.2f ms." % (fl / 1e12, tus / 1e3, fl / tus / 1e6, fl / tus / 1e6 / tf_peak, 3 * fl / tus / 1e6 / tf_peak,
                                                    sum(o["us"] for o in ops.values() if not o["flops"]) / 1e3))


if __name__ == "__main__":
    main()
