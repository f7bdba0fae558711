"""Host-side wrappers: torch CUDA tensors in, libvpf.so (C ABI, include/vpf.h) out.

PyTorch is plumbing only (device memory, streams); every computation below runs in the
hand-written sm_100a kernels.  All calls enqueue on the current torch stream and do not sync.
"""
from __future__ import annotations

import ctypes as C
import functools

import torch

from . import _lib

N_PATCH = 196        # patches per crop at the default crop size (224)
PATCH_K = 768
TOKENS = 197
CROP = 224


def n_patch(crop_size: int = CROP) -> int:
    """Patches per crop: (S/16)^2 for a crop size S that is a multiple of 16 in [16, 224]."""
    s = int(crop_size)
    if s < 16 or s > CROP or s % 16:
        raise ValueError(f"crop size must be a multiple of 16 in [16, {CROP}], got {crop_size}")
    return (s // 16) ** 2


def n_tokens(crop_size: int = CROP) -> int:
    return n_patch(crop_size) + 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Runs `fn` with the CUDA device of its first CUDA-tensor argument current.  libvpf takes the
    launch stream from `torch.cuda.current_stream()` and keys its per-device state on
    `cudaGetDevice()`, so a tracker built on cuda:k must work while another device is current
    (a single process driving several GPUs)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _dtype_code(dtype) -> int:
    if dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    if dtype == torch.float16:
        return _lib.DTYPE_F16
    raise ValueError("the encoder's 16-bit operand type must be torch.bfloat16 or torch.float16")


def _dev(t: torch.Tensor, dtype) -> torch.Tensor:
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype, dtype)
    return t


def _state(t: torch.Tensor) -> torch.Tensor:
    """Particle state: fp32 SoA [6, n], unit stride along particles, any row pitch >= n (the
    multi-GPU tracker keeps it inside its exchange record, whose rows are n_max apart)."""
    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == 6, t.shape
    assert t.stride(1) == 1 and t.stride(0) >= t.shape[1], t.stride()
    return t


# ---- particle filter ---------------------------------------------------------------------------
@_on_device
def init_particles(state, cx, cy, sigma_init, seed, stream_id=0, id_offset=0):
    _state(state)
    n = state.shape[1]
    _lib.check(_lib.load().vpf_init_particles(state.data_ptr(), n, state.stride(0), cx, cy,
                                              sigma_init, seed, stream_id, id_offset, _stream()),
               "vpf_init_particles")


@_on_device
def propagate(state, sigma_pos, sigma_vel, sigma_scale, seed, frame_idx, stream_id=0, id_offset=0):
    _state(state)
    n = state.shape[1]
    _lib.check(_lib.load().vpf_propagate(state.data_ptr(), n, state.stride(0), sigma_pos, sigma_vel,
                                         sigma_scale, seed, frame_idx, stream_id, id_offset,
                                         _stream()), "vpf_propagate")


@_on_device
def frame_advance(frame_idx_dev):
    _lib.check(_lib.load().vpf_frame_advance(frame_idx_dev.data_ptr(), _stream()), "vpf_frame_advance")


@_on_device
def propagate_dev(state, sigma_pos, sigma_vel, sigma_scale, seed, frame_idx_dev, stream_id=0,
                  id_offset=0):
    _state(state)
    _lib.check(_lib.load().vpf_propagate_dev(state.data_ptr(), state.shape[1], state.stride(0),
                                             sigma_pos, sigma_vel, sigma_scale, seed,
                                             frame_idx_dev.data_ptr(), stream_id, id_offset,
                                             _stream()), "vpf_propagate_dev")


@_on_device
def resample_dev(cdf_all, totals, n_local, id_offset, seed, stream_id, frame_idx_dev, state_all,
                 state_out, idx):
    _lib.check(_lib.load().vpf_resample_dev(cdf_all.data_ptr(), totals.data_ptr(), totals.numel(),
                                            n_local, id_offset, seed, stream_id,
                                            frame_idx_dev.data_ptr(), state_all.data_ptr(),
                                            state_out.data_ptr(), state_out.stride(0),
                                            idx.data_ptr(), _stream()), "vpf_resample_dev")


def philox_fill(n, id_offset, c1, c2, c3, seed, device="cuda"):
    out = torch.empty((4, n), dtype=torch.int32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().vpf_philox_fill(out.data_ptr(), n, id_offset, c1, c2, c3, seed,
                                               _stream()), "vpf_philox_fill")
    return out


def normals_fill(n, id_offset, frame_idx, stream_id, seed, device="cuda"):
    out = torch.empty((6, n), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().vpf_normals_fill(out.data_ptr(), n, id_offset, frame_idx, stream_id,
                                                seed, _stream()), "vpf_normals_fill")
    return out


def resample_uniform(seed, frame_idx, stream_id=0) -> float:
    return float(_lib.load().vpf_resample_uniform(seed, frame_idx, stream_id))


# ---- crop --------------------------------------------------------------------------------------
@_on_device
def crop_state(frame, state, w0, h0, out=None, dtype=torch.bfloat16, crop_size=CROP):
    _dev(frame, torch.uint8)
    _state(state)
    H, W, _ = frame.shape
    n = state.shape[1]
    rows = n * n_patch(crop_size)
    if out is None:
        out = torch.empty((rows, PATCH_K), dtype=dtype, device=frame.device)
    assert out.is_contiguous() and out.numel() >= rows * PATCH_K
    _lib.check(_lib.load().vpf_crop_state(frame.data_ptr(), H, W, state.data_ptr(), n,
                                          state.stride(0), w0, h0, out.data_ptr(),
                                          _dtype_code(out.dtype), crop_size, _stream()),
               "vpf_crop_state")
    return out


@_on_device
def crop_boxes(frame, boxes, out=None, dtype=torch.bfloat16, crop_size=CROP):
    _dev(frame, torch.uint8)
    _dev(boxes, torch.float32)
    H, W, _ = frame.shape
    n = boxes.shape[0]
    rows = n * n_patch(crop_size)
    if out is None:
        out = torch.empty((rows, PATCH_K), dtype=dtype, device=frame.device)
    assert out.is_contiguous() and out.numel() >= rows * PATCH_K
    _lib.check(_lib.load().vpf_crop_boxes(frame.data_ptr(), H, W, boxes.data_ptr(), n,
                                          out.data_ptr(), _dtype_code(out.dtype), crop_size,
                                          _stream()), "vpf_crop_boxes")
    return out


# ---- ViT stages (test hooks) ---------------------------------------------------------------------
@_on_device
def gemm_bf16(a, w, bias, mode, out=None, pos=None, patches_per_crop=N_PATCH):
    """C = A[M,K] @ W[N,K]^T with epilogue `mode` (see include/vpf.h)."""
    _dev(a, torch.bfloat16); _dev(w, torch.bfloat16); _dev(bias, torch.float32)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        assert mode in (0, 1)
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    res16 = mode >= 2 and out.dtype == torch.float16        # fp16 residual stream (modes 2, 3)
    _lib.check(_lib.load().vpf_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                         out.data_ptr(), _lib.ptr(pos), M, N, K, mode,
                                         patches_per_crop,
                                         _lib.RESIDUAL_F16 if res16 else _lib.RESIDUAL_F32,
                                         _stream()),
               "vpf_gemm_bf16")
    return out


@_on_device
def layernorm(x, gamma, beta):
    _dev(x, torch.float32)
    rows, D = x.shape
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().vpf_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                         out.data_ptr(), rows, D, _stream()), "vpf_layernorm")
    return out


def ln_stats_parts(D: int) -> int:
    return _lib.load().vpf_ln_stats_parts(int(D))


@_on_device
def rowstats16(x16):
    """(mean, 1/sqrt(var + 1e-6)) per row of an fp16 [rows][D] matrix -> f32 [rows][2]."""
    _dev(x16, torch.float16)
    rows, D = x16.shape
    out = torch.empty((rows, 2), dtype=torch.float32, device=x16.device)
    _lib.check(_lib.load().vpf_rowstats16(x16.data_ptr(), out.data_ptr(), rows, D, _stream()),
               "vpf_rowstats16")
    return out


@_on_device
def gemm_ln(x16, wf, bias_f, stats, gelu=False, out=None, dtype=torch.bfloat16):
    """LayerNorm-folded GEMM: x16 fp16 [M][K] (the residual stream), wf fp16 [N][K] = gamma (.) W
    with each row's mean subtracted, stats f32 [M][2] final (mean, rstd) or [P][M][2] partial (sum x, sum x^2) -> `dtype` [M][N]."""
    _dev(x16, torch.float16)
    M, K = x16.shape
    N = wf.shape[0]
    _dev(wf, torch.float16)
    _dev(bias_f, torch.float32); _dev(stats, torch.float32)
    parts = 0 if stats.dim() == 2 else stats.shape[0]
    assert stats.shape[-2:] == (M, 2)
    if out is None:
        out = torch.empty((M, N), dtype=dtype, device=x16.device)
    assert out.is_contiguous() and out.shape == (M, N)
    _lib.check(_lib.load().vpf_gemm_ln(x16.data_ptr(), wf.data_ptr(), bias_f.data_ptr(),
                                       stats.data_ptr(), parts, out.data_ptr(),
                                       M, N, K, 1 if gelu else 0, _dtype_code(out.dtype), _stream()),
               "vpf_gemm_ln")
    return out


@_on_device
def gemm_residual_stats(a, w, bias, x16, out=None):
    """x16 (fp16 [M][N], in place) += a @ w.T + bias; returns the partial row sums the epilogue
    leaves for a following folded GEMM: f32 [P][M][2] = (sum x, sum x^2) per column group."""
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == w.dtype and a.is_contiguous() and w.is_contiguous()
    _dev(x16, torch.float16); _dev(bias, torch.float32)
    P = ln_stats_parts(N)
    if out is None:
        out = torch.zeros((P, M, 2), dtype=torch.float32, device=a.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (P, M, 2)
    _lib.check(_lib.load().vpf_gemm_residual_stats(a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                                   x16.data_ptr(), out.data_ptr(), M, N, K,
                                                   _dtype_code(a.dtype), _stream()),
               "vpf_gemm_residual_stats")
    return out


@_on_device
def attention(qkv, n_crops, heads, tokens=TOKENS):
    _dev(qkv, torch.bfloat16)
    D = heads * 64
    assert qkv.shape == (n_crops * tokens, 3 * D)
    out = torch.empty((n_crops * tokens, D), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.load().vpf_attention(qkv.data_ptr(), out.data_ptr(), n_crops, heads, tokens,
                                         _stream()), "vpf_attention")
    return out


# ---- likelihood / normalisation / resampling ---------------------------------------------------
@_on_device
def loglik(emb, tmpl, tau, logw, local_max):
    n, D = emb.shape
    _lib.check(_lib.load().vpf_loglik(emb.data_ptr(), D, tmpl.data_ptr(), n, tau, logw.data_ptr(),
                                      local_max.data_ptr(), _stream()), "vpf_loglik")


@_on_device
def sumexp(logw, rank_max, local_sum):
    _lib.check(_lib.load().vpf_sumexp(logw.data_ptr(), logw.numel(), rank_max.data_ptr(),
                                      rank_max.numel(), local_sum.data_ptr(), _stream()),
               "vpf_sumexp")


@_on_device
def weights(logw, rank_max, rank_sum, state, w, cdf_local, total_local, stats, accumulate=True):
    """`accumulate=True`: stats += this shard's five sums (caller zeroes `stats` first);
    False: stats is overwritten (one shard per vector: no zeroing kernel needed)."""
    _lib.check(_lib.load().vpf_weights(logw.data_ptr(), logw.numel(), rank_max.data_ptr(),
                                       rank_sum.data_ptr(), rank_max.numel(), state.data_ptr(),
                                       state.stride(0), w.data_ptr(), cdf_local.data_ptr(),
                                       total_local.data_ptr(), stats.data_ptr(),
                                       1 if accumulate else 0, _stream()),
               "vpf_weights")


@_on_device
def resample(cdf_all, totals, n_local, id_offset, u, state_all, state_out, idx):
    R = totals.numel()
    _lib.check(_lib.load().vpf_resample(cdf_all.data_ptr(), totals.data_ptr(), R, n_local,
                                        id_offset, u, state_all.data_ptr(), state_out.data_ptr(),
                                        state_out.stride(0), idx.data_ptr(), _stream()),
               "vpf_resample")


RECORD_HEADER = 8        # VPF_RECORD_HEADER in include/vpf.h


@_on_device
def resample_records(records, n_max, shard_off, n_local, id_offset, u, state_out, idx,
                     stats_out=None, frame_idx_dev=None, seed=0, stream_id=0):
    """Multi-GPU form: `records` f64 [R][rec_len] is the all-gather's receive buffer, read in
    place (layout in include/vpf.h); `shard_off` int32 [R+1] on the device."""
    R, rec_len = records.shape
    assert records.dtype == torch.float64 and records.is_contiguous()
    assert shard_off.dtype == torch.int32 and shard_off.numel() == R + 1
    _lib.check(_lib.load().vpf_resample_records(
        records.data_ptr(), rec_len, n_max, shard_off.data_ptr(), R, n_local, id_offset,
        0.0 if u is None else float(u), _lib.ptr(frame_idx_dev), seed, stream_id,
        state_out.data_ptr(), state_out.stride(0), idx.data_ptr(), _lib.ptr(stats_out), _stream()),
        "vpf_resample_records")


# ---- ViT encoder handle --------------------------------------------------------------------------
class VitEncoder:
    """Owns the opaque vpf_vit handle, the packed device weights and the workspace tensor."""

    def __init__(self, dim, depth, heads, max_crops, packed_weights, device="cuda",
                 dtype=torch.bfloat16, crop_size=CROP, residual_dtype=torch.float32, ln_fold=None):
        """`ln_fold`: the tensors of `vit_weights.pack_ln_fold` (fp16 residual stream only): the two
        per-block LayerNorms are folded into the qkv / fc1 GEMMs, which then read the stream itself."""
        lib = _lib.load()
        self._lib = lib
        self.dim, self.depth, self.heads, self.max_crops = dim, depth, heads, max_crops
        self.dtype = dtype
        self.crop_size = int(crop_size)
        self.tokens = n_tokens(crop_size)
        if tuple(packed_weights[3].shape) != (self.tokens, dim):
            raise ValueError(f"pos_embed is {tuple(packed_weights[3].shape)}, crop size "
                             f"{crop_size} needs ({self.tokens}, {dim}): pack the weights with "
                             "pack_for_device(..., crop_size=...)")
        wd = {t.dtype for t in packed_weights if t.dtype in (torch.bfloat16, torch.float16)}
        if wd != {dtype}:
            raise ValueError(f"packed GEMM weights are {wd}, encoder operand type is {dtype}")
        if residual_dtype not in (torch.float32, torch.float16):
            raise ValueError("residual_dtype must be torch.float32 (default) or torch.float16")
        self.residual_dtype = residual_dtype
        self.cfg = _lib.VitConfig(dim, depth, heads, max_crops, _dtype_code(dtype), self.crop_size,
                                  _lib.RESIDUAL_F16 if residual_dtype == torch.float16
                                  else _lib.RESIDUAL_F32)
        self._h = C.c_void_p()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.check(lib.vpf_vit_create(C.byref(self._h), C.byref(self.cfg)), "vpf_vit_create")
        self.weights = [t.to(self.device).contiguous() for t in packed_weights]   # keep alive
        arr = (C.c_void_p * len(self.weights))(*[t.data_ptr() for t in self.weights])
        _lib.check(lib.vpf_vit_set_weights(self._h, arr, len(self.weights)), "vpf_vit_set_weights")
        nbytes = lib.vpf_vit_workspace_bytes(C.byref(self.cfg))
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        _lib.check(lib.vpf_vit_set_workspace(self._h, aligned, nbytes), "vpf_vit_set_workspace")
        self.fold_weights = None
        if ln_fold is not None:
            if residual_dtype != torch.float16:
                raise ValueError("ln_fold needs residual_dtype=torch.float16 (the stream is the A operand)")
            if len(ln_fold) != 4 * depth or {t.dtype for t in ln_fold[0::2]} != {torch.float16}:
                raise ValueError("ln_fold must be pack_ln_fold(model, sd): 4 tensors per block, W' in fp16")
            self.fold_weights = [t.to(self.device).contiguous() for t in ln_fold]
            arr = (C.c_void_p * len(self.fold_weights))(*[t.data_ptr() for t in self.fold_weights])
            _lib.check(lib.vpf_vit_set_ln_fold(self._h, arr, len(self.fold_weights)),
                       "vpf_vit_set_ln_fold")
        #: 0 = LayerNorm kernels, 1 = folded with a row-statistics kernel, 2 = folded, statistics
        #: from the residual GEMMs' epilogues (env VPF_LN_FOLD chooses; default 2 when folded)
        self.ln_fold = lib.vpf_vit_ln_fold(self._h)
        self.launches = lib.vpf_vit_launches(self._h)

    def forward(self, patches, n_crops, emb=None):
        assert patches.is_cuda and patches.dtype == self.dtype and patches.is_contiguous()
        if patches.device != self.device:
            raise ValueError(f"patches live on {patches.device}, the encoder on {self.device}")
        if emb is None:
            emb = torch.empty((n_crops, self.dim), dtype=torch.float32, device=patches.device)
        if self.device.index != torch.cuda.current_device():
            with torch.cuda.device(self.device):
                return self._forward(patches, n_crops, emb)
        return self._forward(patches, n_crops, emb)

    def _forward(self, patches, n_crops, emb):
        _lib.check(self._lib.vpf_vit_forward(self._h, patches.data_ptr(), n_crops, emb.data_ptr(),
                                             _stream()), "vpf_vit_forward")
        return emb

    TIMING_KINDS = ("patch_embed", "layernorm", "qkv", "attention", "proj", "fc1", "fc2",
                    "last_block_cls_and_norm")

    def set_timing(self, enable: bool):
        """Bracket every launch of a forward with CUDA events (measurement aid; not capturable)."""
        _lib.check(self._lib.vpf_vit_set_timing(self._h, 1 if enable else 0), "vpf_vit_set_timing")

    def timing(self) -> dict:
        """Device milliseconds of the last forward, summed per kernel kind (waits for it)."""
        buf = (C.c_float * len(self.TIMING_KINDS))()
        _lib.check(self._lib.vpf_vit_timing(self._h, buf, len(self.TIMING_KINDS)), "vpf_vit_timing")
        return dict(zip(self.TIMING_KINDS, [float(v) for v in buf]))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.vpf_vit_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
