"""A small pass over every kernel of the encoder and the particle filter, sized for
`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py` on a machine
where the sanitizer may be used (this pool refuses it; there the hand-fenced canary tests in
tests/test_gpu_cropsize.py and tests/test_gpu_lnfold.py stand in): ViT-Ti, 3 crops, folded and
unfolded LayerNorm, fp16 and fp32 stream, ragged GEMM shapes, two tracker steps.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import Tracker, init_state_dict, ops, pack_for_device, pack_ln_fold  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402
from vitparticlefiltertracker_b200.vit_weights import VIT_CONFIGS  # noqa: E402


def main():
    dev = torch.device("cuda")
    model = "vit_tiny_patch16_224"
    cfg = VIT_CONFIGS[model]
    sd = init_state_dict(model)
    n = 3
    g = torch.Generator().manual_seed(0)
    patches = torch.randn(n * 196, 768, generator=g).to(torch.bfloat16).to(dev)
    outs = []
    for rd, fold in ((torch.float32, None), (torch.float16, None), (torch.float16, pack_ln_fold(model, sd))):
        enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, sd), dev,
                             residual_dtype=rd, ln_fold=fold)
        outs.append(enc.forward(patches, n).cpu())
        print("encoder", rd, "fold" if fold else "plain", "ln_fold mode", enc.ln_fold, "finite",
              bool(torch.isfinite(outs[-1]).all()))
    print("fold vs plain rel", float((outs[2] - outs[1]).norm() / outs[1].norm()))
    # every GEMM epilogue with a ragged M (not a multiple of the 256-row tile) and both K regimes
    M = 333
    x = (torch.randn(M, 192, generator=g)).to(torch.float16).to(dev)
    for K in (192, 3072):
        a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(192, K, generator=g) * 0.03).to(torch.bfloat16).to(dev)
        parts = ops.gemm_residual_stats(a, w, torch.zeros(192, device=dev), x)
        wq = (torch.randn(256, 192, generator=g) * 0.05).to(torch.float16).to(dev)
        o = ops.gemm_ln(x, wq, torch.zeros(256, device=dev), parts, gelu=True)
        print("K", K, "parts", tuple(parts.shape), "out finite", bool(torch.isfinite(o.float()).all()))
    frames, gt = make_clip("box", 224, 224, 3, seed=0, target=(48, 48))
    trk = Tracker(model, 32, seed=0)
    trk.init(frames[0], gt[0])
    for t in (1, 2):
        print(trk.step(frames[t]))
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
