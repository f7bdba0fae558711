"""K1 (crop + resize + normalise + patchify) timed alone: achieved HBM GB/s on its algorithmic
bytes (3*S*S*2 written per particle) for several box sizes.  Usage: python tools/bench_crop.py [n_particles]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    frame = torch.randint(0, 256, (1080, 1920, 3), dtype=torch.uint8, generator=g).to(dev)
    res = {"kernel": "crop_col_kernel", "n": n}
    for S in (224, 64):
        patches = torch.empty(n * (S // 16) ** 2, 768, dtype=torch.bfloat16, device=dev)
        for box in (48, 96, 160, 224, 400):
            st = torch.zeros(6, n)
            st[0] = 960 + torch.randn(n, generator=g) * 30
            st[1] = 540 + torch.randn(n, generator=g) * 30
            st[4] = torch.randn(n, generator=g) * 0.05
            st = st.to(dev)
            ms = timeit(lambda: ops.crop_state(frame, st, float(box), float(box), patches, crop_size=S))
            byts = n * 3 * S * S * 2
            res[f"S{S}_box{box}"] = {"ms": round(ms, 4), "GB/s": round(byts / ms / 1e6, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
