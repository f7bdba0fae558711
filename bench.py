#!/usr/bin/env python
"""Benchmark of the ViT particle-filter tracking hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one tracked frame: Tracker.step(frame) = propagate -> crop -> ViT -> weigh ->
resample for every particle.  Workload at N=1: BASELINE.json configs[2] (the headline
single-GPU config): ViT-B/16, 4096 particles, systematic resampling every frame, synthetic
1080p clip.  With N GPUs the particle set grows to 4096*N, sharded by global particle id (weak
scaling, the shape of configs[3]); NCCL carries only the max / exp-sum / CDF exchange.

value  = particle-updates/s (= particles x tracked frames/s), frames already resident in HBM.
e2e    = the same through the public API with HOST frames: pinned host -> device copy of each
         frame and the device -> host read of the state estimate are inside the timed region.
roofline = the tcgen05 GEMM family (qkv/proj/fc1/fc2), timed live with CUDA events inside a
         sustained loop, against MEASURED_PEAKS.json's sustained bf16 figure.
cpu_baseline = the CPU oracle (a port: the upstream repo is a README only) on a bounded sample.
residual_fp32 / fp16_operands = (N=1 only, informative) the headline workload with the encoder's
         residual stream stored in fp32 instead of the default fp16, and with fp16 instead of bf16
         GEMM operands (8x closer to the fp32 oracle, ~5 % slower under the power cap).
native_resolution = (N=1 only, informative, measured after everything above) the same workload
         at crop size 64 -- SURVEY.md section 8f-2 "native-resolution crops": 17 tokens per
         particle instead of 197, position embedding resampled, oracle revised to match
         (OracleTracker(crop_size=64)).  It is NOT the headline: `value`, `e2e`, `roofline` stay
         on the oracle's default 224-pixel crops.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "particle_updates_per_s"
UNIT = "particle-updates/s"
MODEL = "vit_base_patch16_224"
PARTICLES_PER_GPU = 4096
FRAME_HW = (1080, 1920)
TARGET = (96, 96)
CLIP_FRAMES = 16
DEFAULT_RESIDUAL = "fp16"      # storage type of the residual stream in the headline run
DEFAULT_SHARDS = "equal"       # particle split across GPUs in the headline run (N > 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"],
                    tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        # address the GPU by UUID: with CUDA_VISIBLE_DEVICES set, nvidia-smi's index is not CUDA's
        self.rows, self.proc, self.index = [], None, str(index)
        try:
            u = str(torch.cuda.get_device_properties(index).uuid)
            self.index = u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.index, f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_run(steps, warmup, n_sample, frames, gt):
    """Times the CPU oracle (all host threads) on a bounded sample: `n_sample` particles of the
    same workload.  Returns (particle-updates/s, seconds per step, threads)."""
    from oracle import OracleTracker
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    trk = OracleTracker(MODEL, n_sample, seed=0)
    trk.init(frames[0], gt[0])
    for i in range(warmup):
        trk.step(frames[1 + i % (len(frames) - 1)])
    t0 = time.perf_counter()
    for i in range(steps):
        trk.step(frames[1 + (warmup + i) % (len(frames) - 1)])
    dt = time.perf_counter() - t0
    return n_sample * steps / dt, dt / steps, threads


REF_SAMPLE = 32          # particles the CPU arm tracks per step (of PARTICLES_PER_GPU)


def run_reference(args, emit):
    """--impl reference: the reference's CPU implementation of the path.  The upstream repo is a
    README only, so this is the oracle port (oracle/), with every host thread torch can use.
    Each step is a bounded SAMPLE of the workload: REF_SAMPLE of the 4096 particles through the
    same clip (~0.8 s per step on 16 threads), so --steps / --warmup are honoured as given."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vitparticlefiltertracker_b200.synthetic import make_clip
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    frames, gt = make_clip("box", FRAME_HW[0], FRAME_HW[1], CLIP_FRAMES, seed=0, target=TARGET,
                           velocity=(3.0, 1.0))
    v, sec, threads = cpu_oracle_run(steps, warmup, REF_SAMPLE, frames, gt)
    sample = (f"OracleTracker (CPU PyTorch fp32, {threads} threads): {REF_SAMPLE} of "
              f"{PARTICLES_PER_GPU} particles per step, {steps} timed + {warmup} warm-up frames of "
              f"the same 1080p clip; value = particles tracked per second, i.e. the full "
              f"4096-particle step would take {PARTICLES_PER_GPU / v:.0f} s")
    cfg = workload_config(args.gpus)
    cfg["workload"] += (f" -- CPU arm: a {REF_SAMPLE}-particle sample of it per step "
                        "(per-particle cost is constant, the rate extrapolates linearly)")
    cfg["particles_timed_per_step"] = REF_SAMPLE
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s_at_4096": v / PARTICLES_PER_GPU,
    }
    emit(line)


def workload_config(n_gpus, model=MODEL, ppg=PARTICLES_PER_GPU):
    name = ("BASELINE configs[2]: ViT-B/16 @224" if model == MODEL else
            f"BASELINE configs[3] shape ({model}; --model overrides the headline ViT-B/16)")
    return {"workload": name + f", {ppg} particles per GPU, systematic "
                        "resampling every frame, synthetic 1080p moving-target clip",
            "model": model, "particles_per_gpu": ppg,
            "particles_total": ppg * n_gpus, "frame": "1080x1920x3 uint8",
            "parallelism": f"particle-sharded x{n_gpus}",
            "l2": "per-step working set ~10 GB per GPU >> 126 MB L2 (inputs larger than L2)"}


def source_hash(*names):
    """sha256 over the kernel sources a capture depends on (stamped into profiles/ncu_gemm_*.json
    by tools/summarize_ncu.py; a capture of another build is not reported as this build's)."""
    import hashlib
    h = hashlib.sha256()
    for n in names:
        h.update(open(os.path.join(ROOT, "vitparticlefiltertracker_b200", "csrc", n), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(residual):
    """DRAM bytes of the GEMM family from the newest committed `ncu --set full` capture whose
    source stamp matches the GEMM sources of this build and whose residual-stream type is the
    one being run; (None, why) otherwise."""
    cur = source_hash("gemm.cuh", "sm100.cuh")
    for name in ("ncu_gemm_r02f.json", "ncu_gemm_r02e.json", "ncu_gemm_r02c.json", "ncu_gemm_r02b.json", "ncu_gemm_r02.json",
                 "ncu_gemm_r01.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(prof):
            continue
        pj = json.load(open(prof))
        if pj.get("source_sha16") != cur or pj.get("residual", "fp32") != residual:
            continue
        how = pj.get("traffic_note") or ("dram__bytes_read+write summed over the four launches of one "
                                         "block (ncu --set full)")
        return (pj["gemm_family_traffic_bytes_per_block"],
                f"profiles/{name} (capture of git {pj.get('git_sha', '?')}, source stamp {cur}): {how}; "
                f"algorithmic bytes {pj['gemm_family_algorithmic_bytes_per_block']:.4g}")
    return None, f"no committed ncu capture matches this build's GEMM sources (stamp {cur})"


def gemm_roofline(dev, M, D, peaks, reps=2, x16=False, fold=False):
    """Times every GEMM launch of a 12-block encoder pass (qkv, proj, fc1, fc2 at the real M)
    with CUDA events on the launching stream, inside a sustained loop.  `fold`: the kernels the
    default fp16-stream encoder really launches -- qkv / fc1 with the LayerNorm folded in (A = the
    stream, statistics from partial sums), proj / fc2 leaving those sums in their epilogues."""
    from vitparticlefiltertracker_b200 import ops
    xn = (torch.randn(M, D, device=dev) * 0.5).to(torch.bfloat16)
    big = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    X = torch.zeros(M, D, device=dev, dtype=torch.float16 if x16 else torch.float32)
    shapes = [("qkv", 3 * D, D, 0), ("proj", D, D, 2), ("fc1", 4 * D, D, 1), ("fc2", D, 4 * D, 2)]
    ws = {n: ((torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16),
              torch.zeros(N, device=dev)) for n, N, K, _ in shapes}
    qkv_out = big.view(-1)[: M * 3 * D].view(M, 3 * D)
    if fold:
        X.copy_(torch.randn(M, D, device=dev).to(X.dtype))
        parts = torch.zeros((ops.ln_stats_parts(D), M, 2), dtype=torch.float32, device=dev)
        wf = {n: ws[n][0].to(torch.float16) for n in ("qkv", "fc1")}    # folded weights are fp16

    def one(name, N, K, mode):
        w, b = ws[name]
        a = big if name == "fc2" else xn
        out = X if mode == 2 else (qkv_out if name == "qkv" else big)
        if not fold:
            ops.gemm_bf16(a, w, b, mode, out=out)
        elif mode == 2:
            ops.gemm_residual_stats(a, w, b, X, out=parts)
        else:
            ops.gemm_ln(X, wf[name], b, parts, gelu=(mode == 1), out=out)

    for name, N, K, mode in shapes:          # warm-up
        one(name, N, K, mode)
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        for _blk in range(12):
            for name, N, K, mode in shapes:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                one(name, N, K, mode)
                e.record()
                ev.append((name, N, K, s, e))
    torch.cuda.synchronize()
    per = {}
    for name, N, K, s, e in ev:
        d = per.setdefault(name, {"ms": [], "flops": 2.0 * M * N * K})
        d["ms"].append(s.elapsed_time(e))
    tot_ms = sum(float(np.mean(d["ms"])) for d in per.values())
    tot_fl = sum(d["flops"] for d in per.values())
    detail = {k: {"ms": float(np.mean(d["ms"])),
                  "tflops": d["flops"] / float(np.mean(d["ms"])) / 1e9} for k, d in per.items()}
    achieved = tot_fl / tot_ms / 1e9
    peak = peaks["tflops_sustained"]
    traffic, traffic_src = None, None
    if M == 4096 * 197 and D == 768:                       # captures are of this exact shape
        traffic, traffic_src = ncu_traffic("fp16" if x16 else "fp32")
    return {"bound": "tensor", "kernel": "gemm_bf16_kernel<tcgen05, cta_group::2> (qkv+proj+fc1+fc2 of one block"
                                         + ("; LayerNorm folded into qkv / fc1, row sums from proj / fc2)" if fold else ")"),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_burst": peaks["tflops_burst"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
            "peak_source": peaks["source"] + ", sustained bf16", "traffic": traffic,
            "traffic_source": traffic_src,
            "flops_per_launch_set": tot_fl, "ms_per_launch_set": tot_ms, "detail": detail}, tot_ms


def stage_times(dev, trk, frame, iters=5):
    """Per-stage device times of one step (CUDA events), SURVEY.md section 8d: K1 crop, K2 ViT
    encoder, K3 particle update (motion + likelihood + normalisation), K4 resampling."""
    from vitparticlefiltertracker_b200 import ops

    def t(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    n, xch = trk.n, trk.xch
    total, cdf, _ = xch.local_views()
    st = trk.particles.clone()
    nxt = torch.empty_like(st)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    stats = torch.zeros(5, dtype=torch.float64, device=dev)

    def k3():
        ops.propagate(st, 0.0, 0.0, 0.0, 0, 1)
        ops.loglik(trk.emb, trk.template, trk.tau, trk.logw, xch.max_local)
        ops.sumexp(trk.logw, xch.max_all, xch.sum_local)
        ops.weights(trk.logw, xch.max_all, xch.sum_all, st, trk.weights, cdf, total, stats)

    return {"K1_crop_ms": t(lambda: ops.crop_state(frame, st, trk.w0, trk.h0, trk.patches)),
            "K2_vit_ms": t(lambda: trk.encoder.forward(trk.patches, n, trk.emb)),
            "K3_particle_update_ms": t(k3),
            "K4_resample_ms": t(lambda: ops.resample(cdf.view(1, -1), total, n, 0, 0.37, st, nxt, idx))}


def encoder_breakdown(trk, frames_src, steps=3):
    """In-step device time of the encoder's kernels, summed per kind (vpf_vit_set_timing: an event
    pair around every launch of the real step, so shares and absolutes are those of the throttled
    step, unlike the cold serialised ncu launch list).  Every rank steps (collectives); the caller
    reports rank 0's."""
    trk.encoder.set_timing(True)
    acc = {}
    for i in range(steps + 1):
        trk.step(frames_src[1 + i % (CLIP_FRAMES - 1)])
        t = trk.encoder.timing()
        if i:                                   # the first step settles clocks / caches
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v / steps
    trk.encoder.set_timing(False)
    acc["sum"] = sum(acc.values())
    return acc


def hbm_kernels(dev, trk, frame, peaks, iters=10):
    """K1 (crop) and K3/K4 (particle update + resampling) timed alone with CUDA events: achieved
    HBM GB/s on their algorithmic bytes against the measured copy peak."""
    from vitparticlefiltertracker_b200 import ops
    out = {}
    n = trk.n

    def t(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    ms = t(lambda: ops.crop_state(frame, trk.particles, trk.w0, trk.h0, trk.patches))
    gbs = n * 301056 / ms / 1e6
    out["crop_resize_patchify"] = {"ms": ms, "bytes": n * 301056, "GB/s": gbs,
                                   "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
    xch = trk.xch
    total, cdf, _ = xch.local_views()
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    nxt = torch.empty_like(trk.particles)
    ms = t(lambda: ops.resample(cdf.view(1, -1), total, n, 0, 0.37, trk.particles, nxt, idx))
    gbs = n * 64 / ms / 1e6
    out["resample_gather"] = {"ms": ms, "bytes": n * 64, "GB/s": gbs,
                              "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                              "note": "latency-bound: 262 KB of traffic"}
    return out


def native_resolution_point(dev, frames_dev, frames_host, gt, op_dtype, crop_size=64, steps=20,
                            warmup=4):
    """BASELINE configs[2] at the 'native-resolution' crop size (see the module docstring)."""
    from vitparticlefiltertracker_b200 import Tracker, flops_per_crop
    n = PARTICLES_PER_GPU
    trk = Tracker(MODEL, n, seed=0, device=dev, dtype=op_dtype, crop_size=crop_size)
    trk.init(frames_dev[0], gt[0])
    res = {}
    from vitparticlefiltertracker_b200 import FrameIngest, track
    ingest = FrameIngest(tuple(frames_host[0].shape), device=dev)
    for name, src in (("resident", frames_dev), ("e2e_host_frames", frames_host),
                      ("e2e_host_frames_prefetched", frames_host)):
        for i in range(warmup):
            trk.step(src[1 + i % (CLIP_FRAMES - 1)])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        seq = (src[1 + (warmup + i) % (CLIP_FRAMES - 1)] for i in range(steps))
        if name.endswith("prefetched"):          # pinned double-buffered upload beside the step
            for _ in track(trk, seq, ingest):
                pass
        else:
            for f in seq:
                trk.step(f)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
        res[name] = {"ms_per_step": ms, "frames_per_s": 1e3 / ms,
                     "particle_updates_per_s": n * 1e3 / ms}
    fl = flops_per_crop(MODEL, crop_size)
    res.update({"crop_size": crop_size, "tokens_per_particle": (crop_size // 16) ** 2 + 1,
                "gflop_per_particle_update": fl / 1e9, "steps": steps,
                "step_tflops": n * fl / (res["resident"]["ms_per_step"] / 1e3) / 1e12,
                "note": "informative operating point, not the headline: oracle revision "
                        "OracleTracker(crop_size=64), parity in tests/test_gpu_cropsize.py"})
    return res


def alt_point(dev, frames_dev, gt, op_dtype, res_dtype, note, steps=8, warmup=3):
    """The headline workload at another operating point (operand type / residual-stream type):
    informative, measured in the same run, not the headline."""
    from vitparticlefiltertracker_b200 import Tracker
    n = PARTICLES_PER_GPU
    trk = Tracker(MODEL, n, seed=0, device=dev, dtype=op_dtype, residual_dtype=res_dtype)
    trk.init(frames_dev[0], gt[0])
    for i in range(warmup):
        trk.step(frames_dev[1 + i % (CLIP_FRAMES - 1)])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        trk.step(frames_dev[1 + (warmup + i) % (CLIP_FRAMES - 1)])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    del trk
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "frames_per_s": 1e3 / ms, "particle_updates_per_s": n * 1e3 / ms,
            "steps": steps, "operands": str(op_dtype).replace("torch.", ""),
            "residual": str(res_dtype).replace("torch.", ""), "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=MODEL,
                    help="encoder (default: the headline ViT-B/16; vit_large_patch16_224 with "
                         "--gpus 8 is BASELINE configs[3]: 32768 particles sharded over 8 GPUs)")
    ap.add_argument("--no-native", action="store_true",
                    help="skip the informative crop-size-64 operating point")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"],
                    help="16-bit operand type of the encoder (fp32 accumulation either way)")
    ap.add_argument("--residual", default=DEFAULT_RESIDUAL, choices=["fp32", "fp16"],
                    help="storage type of the encoder's residual stream")
    ap.add_argument("--shards", default=DEFAULT_SHARDS, choices=["equal", "balanced"],
                    help="N > 1: equal particle shards, or shards sized from a 3-step calibration "
                         "so that a slower (power-capped) GPU carries fewer particles")
    ap.add_argument("--graph", action="store_true",
                    help="replay each step as one captured CUDA graph (NCCL collectives included)")
    ap.add_argument("--particles-per-gpu", type=int, default=PARTICLES_PER_GPU)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the informative operating points (fp16 operands, other residual "
                         "type, native resolution) and the CPU baseline")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library writes to fd 1 (the NCCL version
    # banner, for one) is sent to stderr, and the line itself goes out through a private copy
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from vitparticlefiltertracker_b200 import Tracker, flops_per_crop
    from vitparticlefiltertracker_b200.synthetic import iou, make_clip

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    n_gpus = world
    n_total = args.particles_per_gpu * n_gpus
    peaks = load_peaks()

    frames, gt = make_clip("box", FRAME_HW[0], FRAME_HW[1], CLIP_FRAMES, seed=0, target=TARGET,
                           velocity=(3.0, 1.0))
    frames_host = torch.from_numpy(frames).pin_memory()
    frames_dev = torch.from_numpy(frames).to(dev)
    op_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    res_dtype = torch.float16 if args.residual == "fp16" else torch.float32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_tracker(shards=None, graph=False):
        t = Tracker(args.model, n_total, seed=0, device=dev, process_group=pg, dtype=op_dtype,
                    residual_dtype=res_dtype, shards=shards, cuda_graph=graph)
        t.init(frames_dev[0], gt[0])
        return t

    def rank_stats(vals):
        """min / median / max over ranks of a per-rank scalar (all ranks call this)."""
        t = torch.tensor([float(vals)], device=dev, dtype=torch.float64)
        if world > 1:
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            v = [float(x.item()) for x in out]
        else:
            v = [float(t.item())]
        return {"min": min(v), "median": float(np.median(v)), "max": max(v), "per_rank": v}

    shards, calibration = None, None
    if world > 1 and args.shards == "balanced":
        # 8 timed steps (after 3 to reach the power-capped steady state) with equal shards; each
        # rank's encoder-side time (events around propagate, crop and the ViT: no collective
        # inside) sizes its shard for the real run
        cal = make_tracker()
        cal.timing = True
        comp = []
        for i in range(11):
            cal.step(frames_dev[1 + i % (CLIP_FRAMES - 1)])
            torch.cuda.synchronize()
            if i >= 3:
                comp.append(cal.last_timing()[0])
        st = rank_stats(float(np.mean(comp)))
        from vitparticlefiltertracker_b200.exchange import balanced_shards
        shards = balanced_shards(n_total, st["per_rank"], quantum=8)
        calibration = {"compute_ms_equal_shards": st, "shards": shards}
        del cal
        torch.cuda.empty_cache()
        barrier()
    trk = make_tracker(shards, args.graph)

    ious, t_comp, t_xch = [], [], []

    def timed(frames_src, steps, warmup, record=False):
        for i in range(warmup):
            f = 1 + i % (CLIP_FRAMES - 1)
            ious.append(iou(trk.step(frames_src[f]).bbox, gt[f]))
        barrier()
        trk.timing = record and not args.graph
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            f = 1 + (warmup + i) % (CLIP_FRAMES - 1)
            ious.append(iou(trk.step(frames_src[f]).bbox, gt[f]))
            if trk.timing:               # step() has just read the estimate back: events are done
                c, x = trk.last_timing()
                t_comp.append(c)
                t_xch.append(x)
        e.record()
        barrier()
        trk.timing = False
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(frames_dev, args.steps, args.warmup, record=True)
    ms_e2e = timed(frames_host, args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    per_rank = None
    if t_comp:
        # where a multi-GPU step's time goes, per rank: the encoder side (propagate + crop + ViT,
        # no collective) and everything after it (likelihood, three collectives, resampling) --
        # on a fast rank the latter is mostly WAITING for the slowest rank's first all-gather
        per_rank = {"compute_ms": rank_stats(float(np.mean(t_comp))),
                    "exchange_and_wait_ms": rank_stats(float(np.mean(t_xch)))}
        c, x = per_rank["compute_ms"], per_rank["exchange_and_wait_ms"]
        per_rank["collective_floor_ms"] = x["min"]
        per_rank["straggler_spread_ms"] = c["max"] - c["min"]
        per_rank["limiter"] = None if world == 1 else (
            "straggler: the slowest rank's encoder time paces every rank (spread "
            f"{c['max'] - c['min']:.2f} ms vs {x['min']:.2f} ms of collectives + PF kernels)"
            if c["max"] - c["min"] > x["min"] else
            f"exchange: collectives + PF kernels {x['min']:.2f} ms per step exceed the rank "
            f"spread {c['max'] - c['min']:.2f} ms")

    breakdown = encoder_breakdown(trk, frames_dev) if not args.graph else None
    x16 = args.residual == "fp16"
    roof, gemm_ms = gemm_roofline(dev, trk.n * 197, trk.dim, peaks, x16=x16,
                                   fold=trk.encoder.ln_fold == 2) if rank == 0 else (None, 0.0)
    hbm = hbm_kernels(dev, trk, frames_dev[1], peaks) if rank == 0 else None
    stages = stage_times(dev, trk, frames_dev[1]) if rank == 0 and not args.graph else None
    extras = {}
    if n_gpus == 1 and not args.no_extras and not args.no_native and args.model == MODEL \
            and args.particles_per_gpu == PARTICLES_PER_GPU:
        other_res = "fp32" if x16 else "fp16"
        extras["fp16_operands" if args.dtype == "bf16" else "bf16_operands"] = alt_point(
            dev, frames_dev, gt, torch.float16 if args.dtype == "bf16" else torch.bfloat16, res_dtype,
            "same workload, the other 16-bit operand type (fp16: 8x closer to the fp32 oracle, "
            "see BASELINE.md section 5; same tensor-core rate and bytes)")
        extras["residual_" + other_res] = alt_point(
            dev, frames_dev, gt, op_dtype, torch.float32 if x16 else torch.float16,
            f"same workload with the residual stream stored in {other_res}")
        extras["native_resolution"] = native_resolution_point(dev, frames_dev, frames_host, gt, op_dtype)
    barrier()

    tracking = {"mean_iou_vs_ground_truth": float(np.mean(ious[: args.warmup + args.steps])),
                "frames": args.warmup + args.steps,
                "note": "seeded random-init ViT features (no pretrained weights offline); "
                        "consecutive clip frames, free-running filter"}
    if rank == 0:
        ms_step = ms_dev / args.steps
        value = n_total * args.steps / (ms_dev / 1e3)
        e2e_v = n_total * args.steps / (ms_e2e / 1e3)
        fl = flops_per_crop(args.model)
        roof["share_of_step"] = gemm_ms * trk.encoder.depth / ms_step   # upper bound: the last block is CLS-only
        launches_per_step = trk.encoder.launches + 7      # + propagate, crop, loglik, max_reduce,
        line = {                                          #   sumexp, weights, resample
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(workload_config(n_gpus, args.model, args.particles_per_gpu),
                           operands=args.dtype, residual_stream=args.residual,
                           shards=(trk.shards if n_gpus > 1 else None),
                           cuda_graph=bool(args.graph),
                           layernorm=("folded into qkv / fc1 (row sums from the proj / fc2 epilogues)",
                                      "folded into qkv / fc1 (row-statistics kernel)")[trk.encoder.ln_fold % 2]
                           if trk.encoder.ln_fold else "separate kernels"),
            "frames_per_s": args.steps / (ms_dev / 1e3),
            "step_tflops_per_gpu": trk.n * fl / (ms_step / 1e3) / 1e12,
            "step_frac_of_sustained_peak": trk.n * fl / (ms_step / 1e3) / 1e12 / peaks["tflops_sustained"],
            "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(frames_host[0].numel()),
                    "d2h_bytes_per_step": 40},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "hbm_kernels": hbm, "stages_ms": stages,
            "encoder_breakdown_ms": breakdown, "tracking": tracking,
            "clocks": clocks,
        }
        line.update(extras)
        if per_rank is not None:
            line["per_rank"] = per_rank
        if calibration is not None:
            line["calibration"] = calibration
        if n_gpus == 1 and not args.no_cpu_baseline and not args.no_extras and args.model == MODEL:
            v, sec, threads = cpu_oracle_run(2, 1, REF_SAMPLE, frames[:4], gt[:4])
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"OracleTracker (CPU PyTorch fp32): {REF_SAMPLE} of {PARTICLES_PER_GPU} "
                          f"particles, 2 timed frames of the same clip ({sec:.1f} s/frame)"}
        emit(line)
    if world > 1:
        barrier()
        if args.graph:
            # a captured graph holds NCCL kernels of this communicator; skip the teardown under it
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
