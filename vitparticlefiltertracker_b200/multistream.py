"""MultiStreamTracker: S independent video streams on one GPU (BASELINE config 5: "concurrent
synthetic needle-like thin-object video streams x 2048 particles each ... streams sharded across
GPUs, no cross-GPU collectives").  Streams are replicas of the single-stream filter: each keeps
its own particle set, template and Philox stream id; the only thing they share is the ViT
encoder, which embeds the crops of ALL streams in one batched forward (that is where the time
goes).  Across GPUs, streams are sharded round-robin by the caller -- nothing is exchanged.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .exchange import Exchange
from .tracker import State
from .vit_weights import (VIT_CONFIGS, init_state_dict, load_checkpoint, pack_for_device,
                          pack_ln_fold)


class MultiStreamTracker:
    def __init__(self, n_streams, model="vit_base_patch16_224", n_particles=2048, seed=0,
                 device="cuda", tau=0.1, sigma_scale=0.01, state_dict=None, first_stream_id=0,
                 crop_size=224, stream_ids=None, residual_dtype=torch.float16):
        if not torch.cuda.is_available():
            raise RuntimeError("MultiStreamTracker needs a CUDA device: there is no CPU fallback")
        self.S, self.n, self.seed, self.tau = int(n_streams), int(n_particles), int(seed), float(tau)
        self.model, self.device = model, torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.crop_size = int(crop_size)
        self.n_patch = ops.n_patch(self.crop_size)
        self.sigma_scale = float(np.float32(sigma_scale))
        # Philox stream ids: global ids when the streams of a job are sharded over several GPUs
        # (stream s -> rank s mod R), so a stream's trajectory does not depend on where it runs
        self.stream_ids = ([first_stream_id + s for s in range(self.S)] if stream_ids is None
                           else [int(s) for s in stream_ids])
        if len(self.stream_ids) != self.S:
            raise ValueError("stream_ids must list one id per stream")
        cfg = VIT_CONFIGS[model]
        self.dim = cfg["dim"]
        sd = load_checkpoint(state_dict, model) if state_dict is not None else init_state_dict(model)
        self.encoder = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], self.S * self.n,
                                      pack_for_device(model, sd, crop_size=self.crop_size),
                                      device=self.device, crop_size=self.crop_size,
                                      residual_dtype=residual_dtype,
                                      ln_fold=(pack_ln_fold(model, sd) if residual_dtype == torch.float16
                                               else None))
        dev, S, n = self.device, self.S, self.n
        f32 = torch.float32
        self.particles = torch.zeros((S, 6, n), dtype=f32, device=dev)
        self._next = torch.zeros((S, 6, n), dtype=f32, device=dev)
        self.weights = torch.zeros((S, n), dtype=f32, device=dev)
        self.patches = torch.empty((S * n * self.n_patch, ops.PATCH_K), dtype=torch.bfloat16, device=dev)
        self.emb = torch.empty((S * n, self.dim), dtype=f32, device=dev)
        self.logw = torch.empty((S, n), dtype=f32, device=dev)
        self.idx = torch.empty((S, n), dtype=torch.int32, device=dev)
        self.templates = torch.zeros((S, self.dim), dtype=f32, device=dev)
        self.xch = [Exchange(None, n, dev) for _ in range(S)]
        self._stats = torch.zeros((S, 5), dtype=torch.float64, device=dev)
        self.frame_idx = -1

    def _frames(self, frames):
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[0] != self.S or frames.shape[3] != 3:
            raise ValueError("frames must be uint8 [S, H, W, 3]")
        return frames.to(self.device, non_blocking=True).contiguous()

    def init(self, frames, bboxes) -> None:
        frames = self._frames(frames)
        self.w0, self.h0, self.sig = [], [], []
        boxes = torch.empty((self.S, 4), dtype=torch.float32)
        for s, bb in enumerate(bboxes):
            x, y, w, h = [float(v) for v in bb]
            self.w0.append(w); self.h0.append(h)
            sc = math.sqrt(w * h)
            self.sig.append((float(np.float32(0.05 * sc)), float(np.float32(0.02 * sc)),
                             float(np.float32(0.1 * sc))))
            boxes[s] = torch.tensor([x + 0.5 * w, y + 0.5 * h, w, h])
        boxes = boxes.to(self.device)
        P = self.n_patch
        tp = torch.empty((self.S * P, ops.PATCH_K), dtype=torch.bfloat16, device=self.device)
        for s in range(self.S):
            ops.crop_boxes(frames[s], boxes[s:s + 1], tp[s * P:(s + 1) * P], crop_size=self.crop_size)
        self.encoder.forward(tp, self.S, self.templates)
        for s, bb in enumerate(bboxes):
            x, y, w, h = [np.float32(v) for v in bb]
            cx = float(np.float32(x + np.float32(0.5) * w))
            cy = float(np.float32(y + np.float32(0.5) * h))
            ops.init_particles(self.particles[s], cx, cy, self.sig[s][2], self.seed, self.stream_ids[s], 0)
        self.frame_idx = 0

    def step(self, frames):
        """One frame of every stream.  Returns a list of S State tuples."""
        if self.frame_idx < 0:
            raise RuntimeError("call init(frames, bboxes) before step(frames)")
        frames = self._frames(frames)
        self.frame_idx += 1
        t, n, P = self.frame_idx, self.n, self.n_patch
        for s in range(self.S):
            sp, sv, _ = self.sig[s]
            ops.propagate(self.particles[s], sp, sv, self.sigma_scale, self.seed, t, self.stream_ids[s], 0)
            ops.crop_state(frames[s], self.particles[s], self.w0[s], self.h0[s],
                           self.patches[s * n * P:(s + 1) * n * P], crop_size=self.crop_size)
        self.encoder.forward(self.patches, self.S * n, self.emb)        # all streams, one batch
        self._stats.zero_()
        for s in range(self.S):
            x = self.xch[s]
            ops.loglik(self.emb[s * n:(s + 1) * n], self.templates[s], self.tau, self.logw[s], x.max_local)
            ops.sumexp(self.logw[s], x.max_all, x.sum_local)
            total, cdf, _ = x.local_views()
            ops.weights(self.logw[s], x.max_all, x.sum_all, self.particles[s], self.weights[s], cdf,
                        total, self._stats[s])
            u = ops.resample_uniform(self.seed, t, self.stream_ids[s])
            ops.resample(cdf, total, n, 0, u, self.particles[s], self._next[s], self.idx[s])
        self.particles, self._next = self._next, self.particles
        out = []
        for s, st in enumerate(self._stats.tolist()):            # one device->host read
            sw = st[0]
            cx, cy, sc = st[2] / sw, st[3] / sw, st[4] / sw
            w, h = self.w0[s] * math.exp(sc), self.h0[s] * math.exp(sc)
            out.append(State((cx - 0.5 * w, cy - 0.5 * h, w, h), 1.0 / st[1], t))
        return out
