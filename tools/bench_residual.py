"""A/B of the residual-stream storage type (fp32 default vs VPF_RESIDUAL_F16) on BASELINE
configs[2] (ViT-B/16, 4096 particles, 1080p), alternating runs in one process, plus crop size 64."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import Tracker  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


def timed(step, frames, steps, warmup=3):
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
    dev = torch.device("cuda")
    frames, gt = make_clip("box", 1080, 1920, 16, seed=0, target=(96, 96), velocity=(3.0, 1.0))
    fd = torch.from_numpy(frames).to(dev)
    out = {}
    for S, steps in ((224, 8), (64, 40)):
        trk = {rd: Tracker("vit_base_patch16_224", 4096, seed=0, crop_size=S, residual_dtype=rd)
               for rd in (torch.float32, torch.float16)}
        for t in trk.values():
            t.init(fd[0], gt[0])
        res = {"fp32": [], "fp16": []}
        for _ in range(3):
            res["fp32"].append(round(timed(trk[torch.float32].step, fd, steps), 3))
            res["fp16"].append(round(timed(trk[torch.float16].step, fd, steps), 3))
        out[f"crop{S}_ms_per_frame"] = res
        del trk
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
