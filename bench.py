#!/usr/bin/env python
"""Benchmark of the offline Paraformer-large hot path (BASELINE.json metric: RTFx = audio-seconds / second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|bf16x3|bf16x6|bf16] [--impl reference]

One step = one pass of the hot path (fused Fbank+LFR+CMVN -> 50-layer SAN-M encoder -> CIF predictor -> 16-layer
decoder -> greedy ids) over one batch of 64 synthetic 30 s utterances per GPU (BASELINE configs[1]); N>1 shards
utterances over ranks (weak scaling, one all-gather of token ids per step).  Prints ONE JSON line (rank 0).
`value` times the path with the waveforms already resident in HBM; `e2e` times the same work through the plugin call
ParaformerB200.inference with HOST (pinned) waveforms in and token ids out, copies inside the timed region.
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the reference itself is Python
and /root/reference does not exist on the GPU box) on all host threads for the same metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "RTFx (audio-sec/s) Paraformer-large 30s utts"
UTT_SECONDS = 30.0
UTT_SAMPLES = 480000
BATCH = 64
# algorithmic FLOPs per 30 s utterance: encoder 183.2 G + predictor 0.787 G + decoder 8.389 G + 0.1132 G/token
def flops_per_utt(ntok):
    return (183.2 + 0.787 + 8.389 + 0.1132 * ntok) * 1e9


def log(msg):
    print("[bench %6.1fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def usable_cpus() -> int:
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota (os.cpu_count() reports the
    host's cores inside a limited container and oversubscribing OpenMP threads makes the CPU leg crawl)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_batch(rank):
    from funasr_b200 import synth
    base = [synth.make_wav(UTT_SAMPLES, 1000 + 16 * rank + i, "speechlike") for i in range(8)]
    g = torch.Generator().manual_seed(4242 + rank)
    gains = 0.4 + 0.6 * torch.rand(BATCH, generator=g)
    wavs = [(base[i % 8].roll(1601 * i) * gains[i]).contiguous() for i in range(BATCH)]
    return wavs


def cpu_baseline(time_cap_s=240):
    """The oracle (CPU restatement of the reference, kind 'port') on a bounded sample, in a subprocess with a hard time
    cap: this very script's --impl reference leg (2 x 30 s utterances per step, batch 1 like the reference's CPU
    default auto_model.py:785, all usable host threads)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=time_cap_s, text=True,
                           env={**os.environ, "RANK": "0", "WORLD_SIZE": "1", "CUDA_VISIBLE_DEVICES": ""})
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": "reference leg printed no JSON (rc=%d)" % r.returncode}
    except subprocess.TimeoutExpired:
        return {"error": "CPU leg exceeded %d s" % time_cap_s}


def run_reference(args):
    from funasr_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.PARAFORMER_LARGE
    state = synth.make_state_dict(cfg, 0)
    cmvn = synth.make_cmvn(cfg, 1)
    threads = usable_cpus()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import paraformer_oracle as O
    torch.set_num_threads(threads)
    log("reference leg: oracle on %d threads" % threads)
    n_utts = 2
    wavs = [synth.make_wav(UTT_SAMPLES, 1000 + i, "speechlike") for i in range(n_utts)]
    for _ in range(max(1, args.warmup)):
        O.paraformer_forward(wavs[:1], state, cmvn, cfg.enc_layers, cfg.dec_layers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for w in wavs:
            O.paraformer_forward([w], state, cmvn, cfg.enc_layers, cfg.dec_layers)
    dt = time.perf_counter() - t0
    val = args.steps * n_utts * UTT_SECONDS / dt
    sample = "each step = %d x 30 s utterances (batch 1) of the 64-utterance workload" % n_utts
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "audio-sec/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1000, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Paraformer-large, batch=64 synthetic 30 s utterances per GPU (bounded CPU sample)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "audio-sec/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default=os.environ.get("FA_GEMM_MODE", "bf16x3"))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from funasr_b200 import _abi, synth
    from funasr_b200.engine import FrontendEngine, ParaformerEngine
    import funasr_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # NCCL would print its banner to stdout ahead of the one JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load()
    args.warmup = max(args.warmup, 3)

    cfg = synth.PARAFORMER_LARGE
    state = synth.make_state_dict(cfg, 0)              # identical on every rank (seeded)
    cmvn = synth.make_cmvn(cfg, 1)
    fe = FrontendEngine(cmvn, dev)
    eng = ParaformerEngine(state, cfg, dev, gemm_mode=args.mode)
    wavs = make_batch(rank)
    wav_dev = torch.stack(wavs).to(dev)
    lens_dev = torch.full((BATCH,), UTT_SAMPLES, dtype=torch.int32, device=dev)
    gather_buf = torch.empty((world * BATCH, 514), dtype=torch.int32, device=dev) if world > 1 else None

    def step_device():
        feats, fl = fe(wav_dev, lens_dev, 500)
        out = eng.forward_feats(feats, fl)
        if world > 1:                                   # the job's one collective: token ids of every rank
            mine = torch.full((BATCH, 514), -1, dtype=torch.int32, device=dev)
            ids = out["ids_dev"][:, :512]
            mine[:, 2:2 + ids.shape[1]] = ids
            mine[:, 1] = out["ids_lens_dev"]
            dist.all_gather_into_tensor(gather_buf, mine)   # NCCL over NVLink: 64 x 514 int32 per rank
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    log("engine ready (mode %s), warm-up" % args.mode)
    for _ in range(args.warmup):
        out = step_device()
    log("timed region")
    ntok_mean = float(out["token_num"].float().mean())
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = lib.fa_launch_count()
    r0 = getattr(eng, "replayed_launches", 0)       # kernels replayed from the decoder's CUDA graph are not seen by the C-side counter
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    sync_all()
    launches = int(lib.fa_launch_count() - l0) + int(getattr(eng, "replayed_launches", 0) - r0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    log("device-resident: %.2f ms/step" % (ms_total / args.steps))
    # ---- e2e: plugin call, host (pinned) waveforms in -> token ids on host out
    model = funasr_b200.ParaformerB200(
        encoder="SANMEncoderB200", encoder_conf=dict(output_size=512, attention_heads=4, linear_units=2048, num_blocks=cfg.enc_layers,
                                                      input_layer="pe", kernel_size=11, sanm_shfit=0, selfattention_layer_type="sanm"),
        decoder="ParaformerSANMDecoderB200", decoder_conf=dict(attention_heads=4, linear_units=2048, num_blocks=cfg.dec_layers,
                                                                att_layer_num=cfg.dec_layers, kernel_size=11, sanm_shfit=0),
        predictor="CifPredictorV2B200", predictor_conf=dict(idim=512, threshold=1.0, l_order=1, r_order=1, tail_threshold=0.45),
        input_size=560, vocab_size=cfg.vocab, gemm_mode=args.mode)
    model._engine = eng                                  # same packed weights (saves 0.9 GB + repack time)
    model.cfg = cfg
    frontend = funasr_b200.WavFrontendB200(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6,
                                           dither=0.0, cmvn=cmvn)
    frontend._engine = fe
    host_wavs = [w.pin_memory() for w in wavs]
    keys = ["utt%d" % i for i in range(BATCH)]
    for _ in range(2):
        res, meta = model.inference(host_wavs, key=keys, tokenizer=None, frontend=frontend, device=dev)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, meta = model.inference(host_wavs, key=keys, tokenizer=None, frontend=frontend, device=dev)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    n_max = int(out["token_num"].max())
    clocks = sampler.stop() if sampler else None

    log("e2e done: %.2f ms/step" % (e2e_s / args.steps * 1e3))
    if rank == 0:
        pk, pk_src = peaks()
        audio_per_step = world * BATCH * UTT_SECONDS
        value = audio_per_step * args.steps / (ms_total / 1000)
        # ---- roofline of the dominant kernel (the tcgen05 GEMM): FFN w_1 shape of one encoder layer, timed alone with
        #      CUDA events on the launching stream (burst peak applies); algorithmic flops = 2*M*N*K per launch
        roof = None
        try:
            roof = dominant_gemm_roofline(lib, eng, dev, args.mode, pk, pk_src)
        except Exception as e:  # pragma: no cover
            roof = {"error": str(e)}
        line = {"metric": METRIC, "value": value, "unit": "audio-sec/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3->f32", "bf16x6": "bf16x6->f32", "bf16": "bf16"}[args.mode],
                "data": "synthetic",
                "config": {"workload": "Paraformer-large (50 enc + 16 dec layers, vocab 8404, 220 M params, seeded synthetic weights), "
                                       "batch=64 synthetic 30 s 16 kHz utterances per GPU, fused Fbank+encoder+CIF+decoder+greedy",
                           "batch_per_gpu": BATCH, "utt_seconds": UTT_SECONDS, "gemm_mode": args.mode, "tokens_per_utt_mean": ntok_mean,
                           "n_max": n_max, "parallelism": "utterance-sharded dp%d" % world,
                           "l2": "per-step working set (0.9 GB weights + >1 GB activations) exceeds the 126 MB L2; no flush needed",
                           "algorithmic_gflop_per_utt": flops_per_utt(ntok_mean) / 1e9},
                "clocks": clocks,
                "e2e": {"value": audio_per_step * args.steps / e2e_s, "unit": "audio-sec/s", "h2d_bytes_per_step": BATCH * UTT_SAMPLES * 4 + BATCH * 4,
                        "d2h_bytes_per_step": BATCH * 4 + BATCH * n_max * 4 + BATCH * 4, "api": "ParaformerB200.inference(list of pinned host waveforms)"},
                "gpu_launches": launches,
                "achieved_tflops_algorithmic": flops_per_utt(ntok_mean) * world * BATCH * args.steps / (ms_total / 1000) / 1e12,
                "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            try:
                log("cpu baseline leg")
                line["cpu_baseline"] = cpu_baseline()
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dominant_gemm_roofline(lib, eng, dev, mode, pk, pk_src):
    """The dominant kernel = the tcgen05 GEMM (68 % of step time in profiles/r1_launches_bf16x3_b64_v5.csv).  Timed ALONE
    (operand planes pre-split, exactly the launch the encoder makes for FFN w_1: M=32000, N=2048, K=512) with CUDA events
    on the launching stream, L2 flushed between launches; algorithmic flops 2MNK vs the measured bf16 burst peak."""
    import ctypes as C
    from funasr_b200 import _abi
    M, K, N = BATCH * 500, 512, 2048
    lin = eng.enc_layers[1].w1
    x = torch.randn(M, K, device=dev)
    y = torch.empty(M, N, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    passes = {"fp32": 1, "bf16": 1, "bf16x3": 3, "bf16x6": 6}[mode]
    algo = 2.0 * M * N * K
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    gm = _abi.GEMM_MODES[mode]
    npl = {"bf16": 1, "bf16x3": 2, "bf16x6": 3}.get(mode, 0)
    planes = None
    if npl:
        planes = torch.empty(npl, M, K, dtype=torch.bfloat16, device=dev)
        _abi.check(lib.fa_split_rows(x.data_ptr(), K, M, K, K, npl, planes.data_ptr(), st), "fa_split_rows")
    times = []
    for i in range(9):
        flush.zero_()                                      # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if npl:
            _abi.check(lib.fa_linear_planes(planes.data_ptr(), M, C.byref(lin), 1, None, 0, None, 0, y.data_ptr(), N, gm, st), "fa_linear_planes")
        else:
            _abi.check(lib.fa_linear(x.data_ptr(), K, M, C.byref(lin), 1, None, 0, None, 0, y.data_ptr(), N, gm, None, 0, st), "fa_linear")
        e1.record()
        torch.cuda.synchronize(dev)
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    ach = algo / (ms / 1e3) / 1e12
    if mode == "fp32":
        return {"bound": "fp32-simt", "kernel": "gemm_f32_kernel (FFN w_1, M=32000 N=2048 K=512)", "achieved": ach, "peak": None,
                "unit": "TFLOP/s", "frac": None, "traffic": None, "ms": ms}
    peak = pk.get("bf16_tflops", 1590.0)
    return {"bound": "tensor", "kernel": "gemm_tc2_kernel<3,2,EPI_F32> (FFN w_1: M=32000 N=2048 K=512, %s, cta_group::2)" % mode,
            "achieved": ach, "peak": peak, "peak_source": pk_src + " bf16 burst (MEASURED_PEAKS.json)", "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": 278.1e6, "traffic_source": "profiles/r1_ncu_gemm_w1_v6.txt: dram read 69.8 MB + write 208.3 MB per launch "
                                                  "(algorithmic: 69.7 MB planes+weights in, 262 MB fp32 out)",
            "ms": ms, "tensor_passes": passes, "tensor_issue_tflops": ach * passes, "tensor_issue_frac": ach * passes / peak,
            "note": "achieved = ALGORITHMIC fp32-equivalent flops (2MNK) / event time; the bf16x3 split issues 3 bf16 MMAs per "
                    "product for ~2^-17 relative accuracy, so the tensor pipe runs at tensor_issue_frac of the measured peak "
                    "(ncu: sm__pipe_tensor_cycles_active 80 %)"}


if __name__ == "__main__":
    main()
