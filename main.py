#!/usr/bin/env python
"""`python main.py [config.yaml]` -- the entry point the upstream README names (README.md:34-42):
"provide the input data (e.g., video feed or images) and configure the tracking parameters in the
config.yaml file.  The system will output the tracked positions of the objects in real-time."

A thin shim over vitparticlefiltertracker_b200.Tracker: one JSON line per frame on stdout.
Needs a CUDA device (there is no CPU fallback)."""
import json
import sys
import time

import torch
import yaml

from vitparticlefiltertracker_b200 import Tracker, track
from vitparticlefiltertracker_b200.sources import array_frames, open_source
from vitparticlefiltertracker_b200.synthetic import make_clip


def main(path="config.yaml"):
    cfg = yaml.safe_load(open(path))
    inp = cfg.get("input", {})
    kind = inp.get("kind", "synthetic")
    if kind == "synthetic":
        s = inp.get("synthetic", {})
        clip, gt = make_clip(s.get("clip", "box"), s.get("height", 1080), s.get("width", 1920),
                             s.get("frames", 16), s.get("seed", 0), tuple(s.get("target", (96, 96))),
                             velocity=s.get("velocity"))
        frames = array_frames(clip)
        bbox = inp.get("bbox") or gt[0]
    else:                                   # npy | images | video (file or camera index)
        frames = open_source(kind, inp["path"], max_frames=inp.get("max_frames"))
        bbox = inp.get("bbox")
        if bbox is None:
            raise SystemExit(f"input.bbox (x, y, w, h of the target in the first frame) is required for {kind} input")
    sd = cfg.get("checkpoint") or None      # a path: Tracker -> vit_weights.load_checkpoint
    dtype = torch.float16 if cfg.get("dtype", "bf16") == "fp16" else torch.bfloat16
    trk = Tracker(cfg.get("model", "vit_base_patch16_224"), cfg.get("n_particles", 4096),
                  seed=cfg.get("seed", 0), tau=cfg.get("tau", 0.1), sigma_pos=cfg.get("sigma_pos"),
                  sigma_vel=cfg.get("sigma_vel"), sigma_scale=cfg.get("sigma_scale", 0.01),
                  sigma_init=cfg.get("sigma_init"), state_dict=sd, dtype=dtype,
                  cuda_graph=bool(cfg.get("cuda_graph", False)),
                  crop_size=int(cfg.get("crop_size", 224)),
                  residual_dtype=torch.float32 if cfg.get("residual_dtype", "fp16") == "fp32"
                  else torch.float16,
                  ln_fold=cfg.get("ln_fold"))      # None: folded whenever the stream is fp16
    frames = iter(frames)
    trk.init(next(frames), bbox)
    t0 = time.perf_counter()
    for st in track(trk, frames):           # pinned double-buffered upload beside the tracking
        t1 = time.perf_counter()
        print(json.dumps({"frame": st.frame_idx, "bbox": [round(v, 3) for v in st.bbox],
                          "ess": round(st.ess, 2), "ms": round((t1 - t0) * 1e3, 2)}), flush=True)
        t0 = t1


if __name__ == "__main__":
    main(*sys.argv[1:2])
