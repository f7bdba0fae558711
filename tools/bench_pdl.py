"""A/B of Programmatic Dependent Launch on the per-frame kernel chain (run once with VPF_PDL=0 and
once without): ViT-Ti / 256 particles (BASELINE configs[0] shape, CUDA-graph and eager), ViT-B /
4096 particles at crop size 64, and the headline shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import Tracker  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


def timed(step, frames, steps, warmup=4):
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
    out = {"VPF_PDL": os.environ.get("VPF_PDL", "1")}
    frames, gt = make_clip("box", 224, 224, 16, seed=0, target=(48, 48))
    fd = torch.from_numpy(frames).to(dev)
    for graph in (False, True):
        trk = Tracker("vit_tiny_patch16_224", 256, seed=0, cuda_graph=graph)
        trk.init(fd[0], gt[0])
        out[f"vitTi_256_{'graph' if graph else 'eager'}_ms"] = timed(trk.step, fd, 200)
    frames, gt = make_clip("box", 1080, 1920, 16, seed=0, target=(96, 96), velocity=(3.0, 1.0))
    fd = torch.from_numpy(frames).to(dev)
    for graph in (False, True):
        trk = Tracker("vit_base_patch16_224", 4096, seed=0, crop_size=64, cuda_graph=graph)
        trk.init(fd[0], gt[0])
        out[f"vitB_4096_crop64_{'graph' if graph else 'eager'}_ms"] = timed(trk.step, fd, 40)
        del trk
    trk = Tracker("vit_base_patch16_224", 4096, seed=0)
    trk.init(fd[0], gt[0])
    out["vitB_4096_crop224_eager_ms"] = timed(trk.step, fd, 8, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
