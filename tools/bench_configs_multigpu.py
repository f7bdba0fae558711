"""BASELINE.json configs[3] and configs[4] AS NAMED, on N GPUs of one box (run under torchrun):

  cfg3: ViT-L/16, 32 768 particles in total, sharded across 2 / 4 / 8 B200 (strong scaling),
        NCCL global weight normalisation + resampling-CDF exchange
  cfg4: 64 concurrent needle streams x 2048 particles, ViT-B/16, streams round-robin over the
        ranks (8 streams per GPU at 8 GPUs), zero collectives on the data path

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/bench_configs_multigpu.py --config cfg3 [--steps 3 --warmup 2]

Rank 0 prints one JSON line in bench.py's schema (time = max over ranks, CUDA events) and
appends it to gpurun_out/bench_configs_multigpu.jsonl.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vitparticlefiltertracker_b200 import MultiStreamTracker, Tracker, flops_per_crop  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import iou, make_clip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["cfg3", "cfg4"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=64)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, frames):
        nf = len(frames)
        for i in range(args.warmup):
            step(frames[1 + i % (nf - 1)])
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(args.steps):
            step(frames[1 + (args.warmup + i) % (nf - 1)])
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / args.steps

    if args.config == "cfg3":
        model, n_total = "vit_large_patch16_224", 32768
        frames, gt = make_clip("box", 1080, 1920, 8, seed=0, target=(96, 96), velocity=(3.0, 1.0))
        fd = torch.from_numpy(frames).to(dev)
        trk = Tracker(model, n_total, seed=0, device=dev,
                      process_group=dist.group.WORLD if world > 1 else None)
        trk.init(fd[0], gt[0])
        ious = []
        ms = timed(lambda f: ious.append(trk.step(f).bbox), fd)
        units, per_gpu = n_total, trk.n
        cfg = {"workload": "BASELINE configs[3]: ViT-L/16, 32768 particles sharded across "
                           f"{world} B200 (strong scaling), NCCL max / exp-sum / CDF-record exchange, "
                           "1080p synthetic clip", "model": model, "particles_total": n_total,
               "particles_per_gpu": per_gpu, "parallelism": f"particle-sharded x{world}"}
        scaling = "strong"
        extra = {"frames_per_s": 1e3 / ms}
    else:
        model, n, S = "vit_base_patch16_224", 2048, args.streams
        mine = [s for s in range(S) if s % world == rank]         # round-robin, no collectives
        clips = [make_clip("needle", 480, 640, 6, seed=s, target=(3, 120), velocity=(1.5, 0.5))
                 for s in mine]
        frames = torch.from_numpy(np.stack([c[0] for c in clips], axis=1)).to(dev)   # [F,S_local,H,W,3]
        mst = MultiStreamTracker(len(mine), model, n, seed=0, device=dev, stream_ids=mine)
        mst.init(frames[0], [c[1][0] for c in clips])
        last = []
        ms = timed(lambda f: last.append(mst.step(f)), frames)
        units, per_gpu = S * n, len(mine) * n
        cfg = {"workload": f"BASELINE configs[4]: {S} concurrent needle streams x {n} particles, "
                           f"ViT-B/16, streams round-robin over {world} B200 ({len(mine)} per GPU), "
                           "no cross-GPU collectives on the data path, 640x480 synthetic needle clips",
               "model": model, "streams": S, "particles_per_stream": n, "streams_per_gpu": len(mine),
               "parallelism": f"stream-sharded x{world} (replicas)"}
        scaling = "weak"
        extra = {"stream_frames_per_s": S * 1e3 / ms, "rounds_of_all_streams_per_s": 1e3 / ms,
                 "finite_states": bool(all(np.isfinite(st.bbox).all() for st in last[-1]))}
    if rank == 0:
        fl = flops_per_crop(model)
        line = {"metric": "particle_updates_per_s", "value": units * 1e3 / ms,
                "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": cfg, "step_tflops_per_gpu": per_gpu * fl / ms / 1e9,
                "step_frac_of_sustained_peak": per_gpu * fl / ms / 1e9 / peaks["bf16_tflops_sustained"],
                **extra}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_configs_multigpu.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
