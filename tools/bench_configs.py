"""Throughput of the other BASELINE.json configs on ONE B200 (informative; bench.py measures the
headline config).  cfg2: ViT-B, 1024 particles, 720p.  cfg4 (per-GPU share at 8 GPUs): ViT-L, 4096
particles, 1080p.  cfg5 (per-GPU share): 8 needle streams x 2048 particles, ViT-B, 640x480."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import MultiStreamTracker, Tracker, flops_per_crop  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


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


def main():
    out = {}
    dev = torch.device("cuda")
    # cfg2
    frames, gt = make_clip("box", 720, 1280, 12, seed=0, target=(64, 64), velocity=(2.0, 1.0))
    fd = torch.from_numpy(frames).to(dev)
    trk = Tracker("vit_base_patch16_224", 1024, seed=0)
    trk.init(fd[0], gt[0])
    ms = timed(trk.step, fd, 10, 3)
    out["cfg2_vitB_1024_720p"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                                  "particle_updates_per_s": 1024e3 / ms,
                                  "tflops": 1024 * flops_per_crop("vit_base_patch16_224") / ms / 1e9}
    del trk
    torch.cuda.empty_cache()
    # cfg4 per-GPU share
    frames, gt = make_clip("box", 1080, 1920, 8, seed=0, target=(96, 96), velocity=(3.0, 1.0))
    fd = torch.from_numpy(frames).to(dev)
    trk = Tracker("vit_large_patch16_224", 4096, seed=0)
    trk.init(fd[0], gt[0])
    ms = timed(trk.step, fd, 4, 3)
    out["cfg4_vitL_4096_per_gpu"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                                     "particle_updates_per_s": 4096e3 / ms,
                                     "tflops": 4096 * flops_per_crop("vit_large_patch16_224") / ms / 1e9}
    del trk
    torch.cuda.empty_cache()
    # cfg5 per-GPU share: 8 streams x 2048
    S, n = 8, 2048
    clips = [make_clip("needle", 480, 640, 6, seed=s, target=(3, 120), velocity=(1.5, 0.5)) for s in range(S)]
    frames = torch.from_numpy(np.stack([c[0] for c in clips], axis=1)).to(dev)     # [F,S,H,W,3]
    ms_t = MultiStreamTracker(S, "vit_base_patch16_224", n, seed=0)
    ms_t.init(frames[0], [c[1][0] for c in clips])
    ms = timed(ms_t.step, frames, 3, 3)
    out["cfg5_vitB_8x2048_streams_per_gpu"] = {
        "ms_per_round": ms, "stream_frames_per_s": S * 1e3 / ms,
        "particle_updates_per_s": S * n * 1e3 / ms,
        "tflops": S * n * flops_per_crop("vit_base_patch16_224") / ms / 1e9}
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
