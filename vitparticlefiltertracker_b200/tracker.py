"""Tracker: the drop-in Python surface of the hot path.

The upstream repository defines no API; its I/O contract is README.md:42 ("provide the input
data (e.g., video feed or images) ... output the tracked positions").  That is restated as
``Tracker.init(frame, bbox)`` / ``Tracker.step(frame) -> State`` (OpenCV-tracker convention:
bbox = (x, y, w, h), top-left, pixels).  Host code only moves pointers: every stage of a step
runs in libvpf.so's sm_100a kernels on the current CUDA stream -- there is no CPU fallback.

Step order (oracle/tracker.py, D11): propagate -> crop -> ViT -> weigh -> estimate ->
resample -> gather.  ``init`` is frame 0, the first ``step`` is frame 1.

Multi-GPU (``process_group``): particles are sharded by global id (equal shards, or explicit
``shards`` so a slower GPU carries fewer).  Per frame three collectives: all-gather of the local
log-weight maxima, all-gather of the exact local exp-sums, and one all-gather of the packed
record (weight total | estimate sums | local CDF | state) that the resampling kernel reads in
place.  Every quantity exchanged is exact (max, grid-snapped fp64 sums), so the particle
trajectory is bit-identical for any number of ranks and any split.  ``cuda_graph=True`` captures
the whole device side of a frame -- the NCCL collectives included -- into one graph.
"""
from __future__ import annotations

import functools
import math
import os
from typing import NamedTuple

import numpy as np
import torch

from . import ops
from .exchange import Exchange, shard_range, shard_sizes
from .vit_weights import (VIT_CONFIGS, init_state_dict, load_checkpoint, pack_for_device,
                          pack_ln_fold)


class _Stage:
    """NVTX range around one stage of a step (K1 crop, K2 ViT, K3 particle update, K4 resampling) so a
    timeline lines up with SURVEY.md section 8a; active only when VPF_NVTX=1 (no cost otherwise)."""
    enabled = os.environ.get("VPF_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Stage.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Stage.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def _on_own_device(method):
    """Runs a Tracker method with the tracker's device current (a process may drive several
    GPUs; libvpf launches on the current device's stream)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        if self.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.device):
                return method(self, *args, **kwargs)
        return method(self, *args, **kwargs)
    return wrapper


#: the exact fp64 exp-sum of the normalisation (2^-36 grid) needs log2(N) + 36 <= 53
MAX_PARTICLES = 1 << 17


class State(NamedTuple):
    bbox: tuple      # (x, y, w, h) top-left, pixels
    ess: float
    frame_idx: int


class Tracker:
    def __init__(self, model="vit_base_patch16_224", n_particles=4096, seed=0, device="cuda",
                 dtype=torch.bfloat16, tau=0.1, sigma_pos=None, sigma_vel=None, sigma_scale=0.01,
                 sigma_init=None, state_dict=None, process_group=None, stream_id=0,
                 encoder=None, cuda_graph=False, crop_size=224, residual_dtype=torch.float16,
                 shards=None, ln_fold=None):
        if not torch.cuda.is_available():
            raise RuntimeError("Tracker needs a CUDA device: the hot path has no CPU fallback")
        if dtype not in (torch.bfloat16, torch.float16):
            raise ValueError("the sm_100a ViT path takes bf16 (default) or fp16 operands, "
                             "fp32 accumulation")
        self.dtype = dtype
        # Storage type of the encoder's residual stream.  fp16 (default since round 2; saturating
        # conversion, arithmetic stays fp32) halves the residual bytes -- 35 % of a block's HBM
        # traffic -- for -5 % time per frame, and stays inside the stated bf16 tolerance
        # (profiles/parity_r02.json: it adds ~2 % to the bf16 operands' own distance from the
        # oracle).  torch.float32 restores the round-1 behaviour.
        self.residual_dtype = residual_dtype
        # LayerNorm folded into the qkv / fc1 GEMMs (they read the fp16 stream itself; the row
        # statistics come out of the proj / fc2 epilogues): default whenever the stream is fp16.
        self.ln_fold = (residual_dtype == torch.float16) if ln_fold is None else bool(ln_fold)
        if self.ln_fold and residual_dtype != torch.float16:
            raise ValueError("ln_fold needs residual_dtype=torch.float16")
        # 224 = the oracle's default; a smaller multiple of 16 is the "native-resolution crops"
        # option of SURVEY.md section 8f-2 (e.g. 64 -> 17 tokens per particle instead of 197)
        self.crop_size = int(crop_size)
        self.n_patch = ops.n_patch(self.crop_size)
        if model not in VIT_CONFIGS:
            raise ValueError(f"unknown model {model!r}; choose from {sorted(VIT_CONFIGS)}")
        self.model = model
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if not 0 < int(n_particles) <= MAX_PARTICLES:
            raise ValueError(f"n_particles must be in [1, {MAX_PARTICLES}]: the order-independent "
                             "fp64 exp-sum that makes sharded and single-GPU runs bit-identical "
                             "is exact only while log2(N) + 36 <= 53")
        self.seed = int(seed)
        self.tau = float(tau)
        self.stream_id = int(stream_id)
        self._sig = (sigma_pos, sigma_vel, sigma_scale, sigma_init)
        self.pg = process_group
        if self.pg is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        else:
            self.rank, self.world = 0, 1
        self.n_global = int(n_particles)
        self.shards = shard_sizes(self.n_global, self.world, shards)
        self.id_offset, self.n = shard_range(self.n_global, self.rank, self.world, self.shards)
        cfg = VIT_CONFIGS[model]
        self.dim = cfg["dim"]
        if encoder is None:
            sd = (load_checkpoint(state_dict, model) if state_dict is not None
                  else init_state_dict(model))
            encoder = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], self.n,
                                     pack_for_device(model, sd, dtype, self.crop_size),
                                     device=self.device, dtype=dtype, crop_size=self.crop_size,
                                     residual_dtype=residual_dtype,
                                     ln_fold=pack_ln_fold(model, sd, dtype) if self.ln_fold else None)
        assert encoder.max_crops >= self.n and encoder.dtype == dtype
        assert encoder.crop_size == self.crop_size
        self.encoder = encoder
        dev, n = self.device, self.n
        f32, f64 = torch.float32, torch.float64
        self.xch = Exchange(self.pg, n, dev, self.shards)
        if self.world > 1:
            # the rank's particle state lives inside its exchange record (nothing is copied before
            # the all-gather) and resampling writes the children straight back into it
            self.particles = self.xch.state_view()
            self._next = None
        else:
            self.particles = torch.zeros((6, n), dtype=f32, device=dev)
            self._next = torch.zeros((6, n), dtype=f32, device=dev)
        self.weights = torch.full((n,), 1.0 / self.n_global, dtype=f32, device=dev)
        self.patches = torch.empty((n * self.n_patch, ops.PATCH_K), dtype=dtype, device=dev)
        self.emb = torch.empty((n, self.dim), dtype=f32, device=dev)
        self.logw = torch.empty((n,), dtype=f32, device=dev)
        self.idx = torch.empty((n,), dtype=torch.int32, device=dev)
        self.template = torch.zeros((self.dim,), dtype=f32, device=dev)
        self._stats = torch.zeros((5,), dtype=f64, device=dev)
        self._tbox = torch.empty((1, 4), dtype=f32, device=dev)
        self.frame_idx = -1
        self._pending = False
        # CUDA-graph mode: the whole device side of a step (collectives included) is captured once
        # and replayed per frame; the frame counter and the resampling uniform live on the device
        self.use_graph = bool(cuda_graph)
        self._graph = None
        self._frame_buf = None
        self._fidx = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.propagated = None
        # per-frame device timing (bench.py, multi-GPU limiter analysis): events after the encoder
        # and after resampling; off by default
        self.timing = False
        self._ev = None

    # -- helpers -----------------------------------------------------------------------------
    def _to_device_frame(self, frame):
        if isinstance(frame, np.ndarray):
            # e.g. an OpenCV BGR->RGB view (frame[..., ::-1]) has negative strides
            frame = torch.from_numpy(np.ascontiguousarray(frame))
        if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be uint8 [H, W, 3] RGB")
        if not frame.is_cuda:
            frame = frame.to(self.device, non_blocking=True)
        return frame.contiguous()

    # -- API ---------------------------------------------------------------------------------
    @_on_own_device
    def init(self, frame, bbox) -> None:
        frame = self._to_device_frame(frame)
        x, y, w, h = [float(v) for v in bbox]
        if w <= 0 or h <= 0:
            raise ValueError("bbox width and height must be positive")
        self.w0, self.h0 = w, h
        scale = math.sqrt(w * h)
        sp, sv, ss, si = self._sig
        self.sigma_pos = float(np.float32(0.05 * scale if sp is None else sp))
        self.sigma_vel = float(np.float32(0.02 * scale if sv is None else sv))
        self.sigma_scale = float(np.float32(ss))
        self.sigma_init = float(np.float32(0.1 * scale if si is None else si))
        self._tbox.copy_(torch.tensor([[x + 0.5 * w, y + 0.5 * h, w, h]], dtype=torch.float32))
        tp = ops.crop_boxes(frame, self._tbox, dtype=self.dtype, crop_size=self.crop_size)
        self.encoder.forward(tp, 1, self.template.view(1, -1))
        cx = float(np.float32(np.float32(x) + np.float32(0.5) * np.float32(w)))
        cy = float(np.float32(np.float32(y) + np.float32(0.5) * np.float32(h)))
        ops.init_particles(self.particles, cx, cy, self.sigma_init, self.seed, self.stream_id,
                           self.id_offset)
        self.weights.fill_(1.0 / self.n_global)
        self.frame_idx = 0
        self._pending = False
        self._fidx.zero_()
        self._graph = None

    def step(self, frame, _u=None, _emb=None) -> State:
        """One tracked frame.  `_u` (resampling uniform) and `_emb` ([n,D] device embeddings) are
        test hooks that inject oracle quantities; every computation still runs in libvpf."""
        self.step_begin(frame, _u, _emb)
        return self.step_end()

    @_on_own_device
    def step_begin(self, frame, _u=None, _emb=None) -> None:
        """Enqueues the whole device side of one frame on the current stream and returns without
        waiting; `step_end()` reads the estimate back.  Lets a caller do host work (decode, stage the
        next frame: `ingest.track`) while the GPU tracks this one."""
        if self.frame_idx < 0:
            raise RuntimeError("call init(frame, bbox) before step(frame)")
        if self._pending:
            raise RuntimeError("step_begin() called twice without step_end()")
        if self.use_graph and _u is None and _emb is None:
            self._enqueue_graph(frame)
            self._pending = True
            return
        frame = self._to_device_frame(frame)
        self.frame_idx += 1
        if self.use_graph:
            self._fidx.fill_(self.frame_idx)      # keep the device counter in step (hook calls)
        u = ops.resample_uniform(self.seed, self.frame_idx, self.stream_id) if _u is None \
            else float(_u)
        self._device_step(frame, self.frame_idx, u, _emb)
        self.last_u = u
        self._pending = True

    def _device_step(self, frame, t, u, emb_in=None):
        """Device side of one frame.  `t` / `u` are host scalars on the eager path; both None means
        the captured-graph form: static pointers, frame counter and resampling uniform taken from
        the device counter `_fidx`, no host-supplied per-frame values."""
        xch, st, on_dev = self.xch, self.particles, t is None
        if self.timing:
            self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            self._ev[0].record()
        # a3 motion
        with _Stage("K3a propagate"):
            if on_dev:
                ops.frame_advance(self._fidx)
                ops.propagate_dev(st, self.sigma_pos, self.sigma_vel, self.sigma_scale, self.seed,
                                  self._fidx, self.stream_id, self.id_offset)
            else:
                ops.propagate(st, self.sigma_pos, self.sigma_vel, self.sigma_scale, self.seed, t,
                              self.stream_id, self.id_offset)
        # a4 crop, a5 ViT
        if emb_in is None:
            with _Stage("K1 crop"):
                ops.crop_state(frame, st, self.w0, self.h0, self.patches, crop_size=self.crop_size)
            with _Stage("K2 ViT encoder"):
                self.encoder.forward(self.patches, self.n, self.emb)
            emb = self.emb
        else:
            emb = emb_in
        if self.timing:
            self._ev[1].record()
        # a6 likelihood + exact normalisation (collectives are no-ops on a single GPU)
        fdev = self._fidx if on_dev else None
        with _Stage("K3b weigh + normalise"):
            ops.loglik(emb, self.template, self.tau, self.logw, xch.max_local)
            max_all = xch.gather_max()
            ops.sumexp(self.logw, max_all, xch.sum_local)
            sum_all = xch.gather_sum()
            total, cdf, _ = xch.local_views()
            stats = xch.stats_view() if self.world > 1 else self._stats
            ops.weights(self.logw, max_all, sum_all, st, self.weights, cdf, total, stats,
                        accumulate=False)
        # a7 + a8 resample + gather
        with _Stage("K4 resample + gather"):
            if self.world > 1:
                # one all-gather of the packed records; the kernel reads the receive buffer in
                # place, sums the ranks' estimate terms and writes this rank's children back into
                # its own record (the gather has already consumed it)
                rec_all = xch.gather_records()
                ops.resample_records(rec_all, xch.n_max, xch.shard_off, self.n, self.id_offset, u,
                                     st, self.idx, stats_out=self._stats, frame_idx_dev=fdev,
                                     seed=self.seed, stream_id=self.stream_id)
                self.propagated = None
            elif on_dev:
                ops.resample_dev(cdf, total, self.n, self.id_offset, self.seed, self.stream_id,
                                 self._fidx, st, self._next, self.idx)
                st.copy_(self._next)               # keeps every pointer of the graph static
            else:
                ops.resample(cdf, total, self.n, self.id_offset, u, st, self._next, self.idx)
                if self.use_graph:
                    # a captured graph holds these pointers: keep them static
                    st.copy_(self._next)
                    self.propagated = None
                else:
                    self.particles, self._next = self._next, st
                    self.propagated = st           # pre-resampling state of this frame (tests)
        if self.timing:
            self._ev[2].record()

    def last_timing(self):
        """(encoder side ms, exchange + resampling ms) of the last eager frame; needs
        `timing = True` and a synchronised device."""
        e = self._ev
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])

    @_on_own_device
    def step_end(self) -> State:
        """Estimate of the frame enqueued by `step_begin`: the only device->host read of a step."""
        if not self._pending:
            raise RuntimeError("step_end() without step_begin()")
        self._pending = False
        s = self._stats.tolist()
        sw = s[0]
        cx, cy, sc = s[2] / sw, s[3] / sw, s[4] / sw
        ess = 1.0 / s[1] if s[1] > 0 else 0.0
        w = self.w0 * math.exp(sc)
        h = self.h0 * math.exp(sc)
        return State((cx - 0.5 * w, cy - 0.5 * h, w, h), ess, self.frame_idx)

    # -- checkpoint / resume -------------------------------------------------------------------
    def state_dict(self) -> dict:
        """Everything needed to resume this stream: particle set, template, frame counter and
        the scalars fixed by init() (device tensors are cloned)."""
        if self.frame_idx < 0:
            raise RuntimeError("nothing to save before init()")
        return {"particles": self.particles.clone(), "weights": self.weights.clone(),
                "template": self.template.clone(), "frame_idx": self.frame_idx,
                "seed": self.seed, "stream_id": self.stream_id, "w0": self.w0, "h0": self.h0,
                "sigmas": (self.sigma_pos, self.sigma_vel, self.sigma_scale, self.sigma_init),
                "model": self.model, "n_particles": self.n_global, "rank": self.rank,
                "world": self.world, "crop_size": self.crop_size}

    def load_state_dict(self, sd: dict) -> None:
        if (sd["model"], sd["n_particles"], sd["rank"], sd["world"], sd.get("crop_size", 224)) != \
                (self.model, self.n_global, self.rank, self.world, self.crop_size):
            raise ValueError("checkpoint was taken with a different model / particle count / "
                             "sharding / crop size")
        self.particles.copy_(sd["particles"])
        self.weights.copy_(sd["weights"])
        self.template.copy_(sd["template"])
        self.frame_idx, self.seed, self.stream_id = sd["frame_idx"], sd["seed"], sd["stream_id"]
        self.w0, self.h0 = sd["w0"], sd["h0"]
        self.sigma_pos, self.sigma_vel, self.sigma_scale, self.sigma_init = sd["sigmas"]
        self._fidx.fill_(self.frame_idx)
        self._graph = None

    # -- CUDA-graph path ----------------------------------------------------------------------
    def _enqueue_graph(self, frame) -> None:
        if isinstance(frame, np.ndarray):
            frame = torch.from_numpy(np.ascontiguousarray(frame))   # e.g. a BGR->RGB flipped view
        if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be uint8 [H, W, 3] RGB")
        if self._frame_buf is None or self._frame_buf.shape != frame.shape:
            self._frame_buf = torch.empty(frame.shape, dtype=torch.uint8, device=self.device)
            self._graph = None
        self._frame_buf.copy_(frame, non_blocking=True)
        self.frame_idx += 1
        if self._graph is None:
            # first frame eagerly (sets kernel attributes, builds tensor maps, brings the NCCL
            # communicator up); then capture for replay
            timing, self.timing = self.timing, False
            self._device_step(self._frame_buf, None, None)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._device_step(self._frame_buf, None, None)
            # capturing does not execute: the device counter has advanced once (the eager frame)
            self._graph = g
            self.timing = timing
        else:
            self._graph.replay()
        self.last_u = ops.resample_uniform(self.seed, self.frame_idx, self.stream_id)
