"""Throughput of BASELINE configs[2] (ViT-B/16, 4096 particles, 1080p clip, resampling every frame)
as a function of the crop size S ("native-resolution crops").  S = 224 is
the oracle default bench.py measures; S = 64 is the crop size BASELINE configs[1] names as the
native resolution of the target ("64x64 crops upsampled to 224").  Informative: one B200, frames
resident in HBM, CUDA events; eager and CUDA-graph mode.

Usage: python tools/bench_cropsize.py [--sizes 64,96,128,224] [--particles 4096] [--steps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import Tracker, flops_per_crop  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402

MODEL = "vit_base_patch16_224"


def timed(step, frames, steps, warmup):
    nf = len(frames)
    for i in range(warmup):
        step(frames[1 + i % (nf - 1)])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step(frames[1 + (warmup + i) % (nf - 1)])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def iou(a, b):
    ax0, ay0, aw, ah = [float(v) for v in a]
    bx0, by0, bw, bh = [float(v) for v in b]
    iw = max(0.0, min(ax0 + aw, bx0 + bw) - max(ax0, bx0))
    ih = max(0.0, min(ay0 + ah, by0 + bh) - max(ay0, by0))
    inter = iw * ih
    return inter / (aw * ah + bw * bh - inter)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,96,128,224")
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--out", default="gpurun_out/bench_cropsize.json")
    a = ap.parse_args()
    dev = torch.device("cuda")
    n = a.particles
    frames, gt = make_clip("box", 1080, 1920, 16, seed=0, target=(96, 96), velocity=(3.0, 1.0))
    fd = torch.from_numpy(frames).to(dev)
    out = {"model": a.model, "particles": n, "frame": "1080x1920x3", "steps": a.steps, "sizes": {}}
    for S in [int(s) for s in a.sizes.split(",")]:
        res = {"tokens": (S // 16) ** 2 + 1, "gflop_per_particle": flops_per_crop(a.model, S) / 1e9}
        for mode in ("eager", "cuda_graph"):
            steps = a.steps if S < 224 else max(4, a.steps // 4)
            trk = Tracker(a.model, n, seed=0, crop_size=S, cuda_graph=(mode == "cuda_graph"))
            trk.init(fd[0], gt[0])
            ms = timed(trk.step, fd, steps, a.warmup)
            res[mode] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                         "particle_updates_per_s": n * 1e3 / ms,
                         "tflops": n * flops_per_crop(a.model, S) / ms / 1e9}
            if mode == "eager":           # tracking quality over a fresh run of the clip
                trk.init(fd[0], gt[0])
                ious = [iou(trk.step(fd[t]).bbox, gt[t]) for t in range(1, len(frames))]
                res["mean_iou_vs_ground_truth"] = sum(ious) / len(ious)
            del trk
            torch.cuda.empty_cache()
        out["sizes"][str(S)] = res
        print(S, json.dumps(res), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
