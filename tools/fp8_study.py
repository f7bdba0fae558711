"""SURVEY.md section 8f-2, "FP8 GEMMs": decided by measurement, on the CPU, with the oracle.

Emulates reduced-precision GEMM operands (fp32 accumulation) inside the oracle's ViT forward
(oracle/vit.py `quantise`): bf16 and fp16 (to validate the emulation against what the GPU
kernels actually deliver) and e4m3 with per-tensor scaling (what a tcgen05 `kind::f8f6f4` path
would compute), for the linear layers only and for linear + attention contractions.  Reports the
figures the tracker's tolerance is stated in: embedding rel-L2, |dcos| against the template and
the relative error of the normalised weights at tau = 0.1.

    python tools/fp8_study.py [--models vit_tiny_patch16_224,vit_base_patch16_224] [--n 32]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pf  # noqa: E402
from oracle.crop import boxes_from_state, crop_resize  # noqa: E402
from oracle.vit import init_vit_weights, vit_forward  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="vit_tiny_patch16_224,vit_base_patch16_224")
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "fp8_study_r02.json"))
    args = ap.parse_args()
    frames, gt = make_clip("box", 224, 224, 2, seed=0, target=(48, 48))
    x, y, w, h = [float(v) for v in gt[0]]
    st = pf.init_particles(args.n, gt[0], 0.1 * np.sqrt(w * h), seed=0)
    st = pf.propagate(st, 2.4, 0.96, 0.01, 1, 0)
    crops = crop_resize(frames[1], boxes_from_state(st, w, h))
    tcrop = crop_resize(frames[0], np.array([[x + w / 2, y + h / 2, w, h]], np.float32))
    variants = [("bf16 linear + attention (emulates the shipped kernels)", "bf16", "bf16"),
                ("fp16 linear + attention (emulates dtype=float16)", "fp16", "fp16"),
                ("e4m3 linear, bf16 attention", "e4m3", "bf16"),
                ("e4m3 linear + attention", "e4m3", "e4m3")]
    out = {"n_particles": args.n, "tau": 0.1, "rows": []}
    for model in args.models.split(","):
        sd = init_vit_weights(model)
        t0 = time.time()
        ref = vit_forward(crops, sd, model)
        tref = vit_forward(tcrop, sd, model)[0]
        want = pf.weigh(ref, tref, 0.1)
        for name, lf, af in variants:
            emb = vit_forward(crops, sd, model, linear_fmt=lf, attn_fmt=af)
            tmpl = vit_forward(tcrop, sd, model, linear_fmt=lf, attn_fmt=af)[0]
            got = pf.weigh(emb, tmpl, 0.1)
            row = dict(model=model, variant=name,
                       emb_rel_l2=float(np.linalg.norm(emb - ref) / np.linalg.norm(ref)),
                       dcos_max=float(np.abs(got["cos"] - want["cos"]).max()),
                       weights_rel_max=float(np.abs(got["w"] / want["w"] - 1).max()),
                       ess_rel=float(abs(got["ess"] / want["ess"] - 1)))
            out["rows"].append(row)
            print(json.dumps(row), flush=True)
        print(f"# {model}: {time.time() - t0:.0f} s", flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
