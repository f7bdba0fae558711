"""OracleTracker: the whole hot path in CPU fp32 -- test infrastructure / CPU baseline only.

Same surface as ``vitparticlefiltertracker_b200.Tracker``: the reference's only
I/O contract is /root/reference/README.md:42 -- "provide the input data (video feed or images)
... output the tracked positions" -- restated as ``init(frame, bbox)`` / ``step(frame) -> State``.
Step order (D11): propagate -> crop -> ViT -> weigh -> estimate -> resample -> gather;
``init`` is frame_idx 0, the first ``step`` is frame_idx 1.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np

from . import pf
from .crop import boxes_from_state, crop_resize
from .philox import philox_uniform53
from .vit import init_vit_weights, normalise_state_dict, vit_forward


class State(NamedTuple):
    bbox: tuple      # (x, y, w, h) top-left, pixels
    ess: float
    frame_idx: int


class OracleTracker:
    def __init__(self, model="vit_base_patch16_224", n_particles=4096, seed=0, tau=0.1,
                 sigma_pos=None, sigma_vel=None, sigma_scale=0.01, sigma_init=None,
                 state_dict=None, stream_id=0, vit_chunk=32, crop_size=224):
        self.model = model
        self.n = int(n_particles)
        if not 0 < self.n <= pf.MAX_PARTICLES:
            raise ValueError(f"n_particles must be in [1, {pf.MAX_PARTICLES}]: the exact fp64 "
                             "exp-sum (D10) needs log2(N) + 36 <= 53")
        self.seed = int(seed)
        self.tau = float(tau)
        self.stream_id = int(stream_id)
        self._sig = (sigma_pos, sigma_vel, sigma_scale, sigma_init)
        self.sd = (normalise_state_dict(state_dict, model) if state_dict is not None
                   else init_vit_weights(model))
        self.vit_chunk = vit_chunk
        self.crop_size = int(crop_size)     # 224, or a smaller multiple of 16
        self.frame_idx = -1
        self.last = {}

    # -- helpers -----------------------------------------------------------------------------
    def embed(self, frame, boxes):
        crops = crop_resize(np.asarray(frame), boxes, size=self.crop_size)
        return vit_forward(crops, self.sd, self.model, chunk=self.vit_chunk)

    # -- API ---------------------------------------------------------------------------------
    def init(self, frame, bbox):
        x, y, w, h = [float(v) for v in bbox]
        self.w0, self.h0 = np.float32(w), np.float32(h)
        scale = math.sqrt(w * h)
        sp, sv, ss, si = self._sig
        self.sigma_pos = np.float32(0.05 * scale if sp is None else sp)
        self.sigma_vel = np.float32(0.02 * scale if sv is None else sv)
        self.sigma_scale = np.float32(ss)
        self.sigma_init = np.float32(0.1 * scale if si is None else si)
        tbox = np.array([[x + 0.5 * w, y + 0.5 * h, w, h]], dtype=np.float32)
        self.template = self.embed(frame, tbox)[0]
        self.particles = pf.init_particles(self.n, bbox, self.sigma_init, self.seed,
                                           self.stream_id)
        self.weights = np.full(self.n, 1.0 / self.n, dtype=np.float32)
        self.frame_idx = 0

    def step(self, frame, _noise=None, _u=None, _emb=None) -> State:
        assert self.frame_idx >= 0, "call init() first"
        self.frame_idx += 1
        t = self.frame_idx
        st = pf.propagate(self.particles, self.sigma_pos, self.sigma_vel, self.sigma_scale,
                          t, self.seed, self.stream_id, noise=_noise)
        boxes = boxes_from_state(st, self.w0, self.h0)
        emb = self.embed(frame, boxes) if _emb is None else _emb
        res = pf.weigh(emb, self.template, self.tau, st)
        u = philox_uniform53(t, self.stream_id, self.seed) if _u is None else float(_u)
        idx = pf.systematic_resample(res["w"], u)
        self.last = dict(propagated=st, boxes=boxes, emb=emb, u=u, idx=idx, **res)
        self.particles = pf.gather_state(st, idx)
        self.weights = res["w"]
        cx, cy, s = [float(v) for v in res["est"]]
        w = float(self.w0) * math.exp(s)
        h = float(self.h0) * math.exp(s)
        return State((cx - 0.5 * w, cy - 0.5 * h, w, h), res["ess"], t)
