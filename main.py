#!/usr/bin/env python
"""`python main.py [config.yaml]` -- the entry point the upstream README names (README.md:34-42):
"provide the input data (e.g., video feed or images) and configure the tracking parameters in the
config.yaml file.  The system will output the tracked positions of the objects in real-time."

A thin shim over vitparticlefiltertracker_b200.Tracker: one JSON line per frame on stdout.
Needs a CUDA device (there is no CPU fallback)."""
import json
import sys
import time

import numpy as np
import torch
import yaml

from vitparticlefiltertracker_b200 import Tracker
from vitparticlefiltertracker_b200.synthetic import make_clip


def main(path="config.yaml"):
    cfg = yaml.safe_load(open(path))
    inp = cfg.get("input", {})
    if inp.get("kind", "synthetic") == "synthetic":
        s = inp.get("synthetic", {})
        frames, gt = make_clip(s.get("clip", "box"), s.get("height", 1080), s.get("width", 1920),
                               s.get("frames", 16), s.get("seed", 0), tuple(s.get("target", (96, 96))),
                               velocity=s.get("velocity"))
        bbox = inp.get("bbox") or gt[0]
    else:
        frames = np.load(inp["path"], mmap_mode="r")
        bbox = inp["bbox"]
        if bbox is None:
            raise SystemExit("input.bbox is required for npy input")
    sd = torch.load(cfg["checkpoint"], map_location="cpu") if cfg.get("checkpoint") else None
    dtype = torch.float16 if cfg.get("dtype", "bf16") == "fp16" else torch.bfloat16
    trk = Tracker(cfg.get("model", "vit_base_patch16_224"), cfg.get("n_particles", 4096),
                  seed=cfg.get("seed", 0), tau=cfg.get("tau", 0.1), sigma_pos=cfg.get("sigma_pos"),
                  sigma_vel=cfg.get("sigma_vel"), sigma_scale=cfg.get("sigma_scale", 0.01),
                  sigma_init=cfg.get("sigma_init"), state_dict=sd, dtype=dtype,
                  cuda_graph=bool(cfg.get("cuda_graph", False)),
                  crop_size=int(cfg.get("crop_size", 224)),
                  residual_dtype=torch.float16 if cfg.get("residual_dtype", "fp32") == "fp16"
                  else torch.float32)
    trk.init(np.ascontiguousarray(frames[0]), bbox)
    for f in range(1, len(frames)):
        t0 = time.perf_counter()
        st = trk.step(np.ascontiguousarray(frames[f]))
        print(json.dumps({"frame": st.frame_idx, "bbox": [round(v, 3) for v in st.bbox],
                          "ess": round(st.ess, 2), "ms": round((time.perf_counter() - t0) * 1e3, 2)}),
              flush=True)


if __name__ == "__main__":
    main(*sys.argv[1:2])
