"""Two steps of the headline workload (ViT-B/16, 4096 particles, 1080p, 96-pixel target) for
`ncu -k regex:...` captures of individual kernels at the bench shape:

    python tools/prof_kernels.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:crop_col -s 1 -c 1 \
        -o gpurun_out/prof_crop python tools/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import Tracker  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


def main():
    n = int(os.environ.get("VPF_PROF_PARTICLES", "4096"))
    crop = int(os.environ.get("VPF_PROF_CROP_SIZE", "224"))
    frames, gt = make_clip("box", 1080, 1920, 4, seed=0, target=(96, 96), velocity=(3.0, 1.0))
    fd = torch.from_numpy(frames).cuda()
    trk = Tracker("vit_base_patch16_224", n, seed=0, crop_size=crop)
    trk.init(fd[0], gt[0])
    for t in (1, 2):
        print(trk.step(fd[t]))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
