"""Particle-filter arithmetic of the oracle (numpy) -- test infrastructure only.

Restates SURVEY.md Appendix A D5-D10 (anchor: /root/reference/README.md:8, "Particle Filter:
Utilizes probabilistic algorithms for accurate state estimation"; the reference ships no code).

State layout (D5): float32 SoA ``[6, N]`` -- rows cx, cy, vx, vy, s (log-scale), spare.
"""
from __future__ import annotations

import numpy as np

from .philox import philox_normals

CX, CY, VX, VY, S, SPARE = range(6)

#: normalised weights are snapped to multiples of 2^-GRID_BITS before the fp64 scan (D6): every
#: partial sum is then an integer multiple of 2^-40 below 2, so fp64 addition is exact and
#: associative and the CDF does not depend on scan topology or rank count.
GRID_BITS = 40
#: exp(l_i - max) terms are snapped to multiples of 2^-SUM_GRID_BITS before the fp64 sum that
#: feeds the log-sum-exp; N <= 2^17 terms <= 1 -> every partial sum is an integer multiple of
#: 2^-36 not above 2^17, i.e. fits the 53-bit significand: exact.  MAX_PARTICLES enforces it.
SUM_GRID_BITS = 36
MAX_PARTICLES = 1 << (53 - SUM_GRID_BITS)


def init_particles(n: int, bbox, sigma_init: float, seed: int, stream_id: int = 0,
                   id_offset: int = 0) -> np.ndarray:
    """Particle cloud at frame 0: centre + sigma_init * N(0,1) (draw 0, rows z0,z1), v = 0, s = 0."""
    x, y, w, h = [np.float32(v) for v in bbox]
    ids = np.arange(id_offset, id_offset + n, dtype=np.uint32)
    z = philox_normals(ids, 0, stream_id, seed)
    st = np.zeros((6, n), dtype=np.float32)
    cx0 = np.float32(x + np.float32(0.5) * w)
    cy0 = np.float32(y + np.float32(0.5) * h)
    st[CX] = cx0 + np.float32(sigma_init) * z[0]
    st[CY] = cy0 + np.float32(sigma_init) * z[1]
    return st


def propagate(state: np.ndarray, sigma_pos: float, sigma_vel: float, sigma_scale: float,
              frame_idx: int, seed: int, stream_id: int = 0, id_offset: int = 0,
              noise: np.ndarray | None = None) -> np.ndarray:
    """Constant-velocity Gaussian motion, dt = 1 frame (D8).  Returns a new [6,N] fp32 array.

    cx += vx + sp*e0 ; cy += vy + sp*e1 ; vx += sv*e2 ; vy += sv*e3 ; s += ss*e4
    (positions use the velocity from *before* this step's velocity noise).
    """
    n = state.shape[1]
    if noise is None:
        ids = np.arange(id_offset, id_offset + n, dtype=np.uint32)
        noise = philox_normals(ids, frame_idx, stream_id, seed)
    z = noise.astype(np.float32)
    sp, sv, ss = np.float32(sigma_pos), np.float32(sigma_vel), np.float32(sigma_scale)
    out = state.astype(np.float32).copy()
    out[CX] = (state[CX] + state[VX]) + sp * z[0]
    out[CY] = (state[CY] + state[VY]) + sp * z[1]
    out[VX] = state[VX] + sv * z[2]
    out[VY] = state[VY] + sv * z[3]
    out[S] = state[S] + ss * z[4]
    return out


def weigh(emb: np.ndarray, tmpl: np.ndarray, tau: float, state: np.ndarray | None = None):
    """Cosine likelihood, log-sum-exp normalisation, ESS and weighted-mean estimate (D9, D10).

    emb [N, D] fp32, tmpl [D] fp32.  Returns dict(cos, logw, w, lse, ess, est) with w fp32 [N].
    Dot products and norms accumulate in fp64 and the cosine is rounded once to fp32; the GPU
    does the same (fp64 accumulation, any order: the fp64 sums differ by ~1e-16 relative, which
    the single rounding to fp32 absorbs), so log-weights agree bit for bit given the embeddings.
    """
    e = emb.astype(np.float64)
    t = tmpl.astype(np.float64)
    dot = e @ t
    ne = np.maximum(np.sqrt((e * e).sum(axis=1)), 1e-8)
    nt = max(float(np.sqrt((t * t).sum())), 1e-8)
    cos = (dot / (ne * nt)).astype(np.float32)
    logw = (cos / np.float32(tau)).astype(np.float32)
    out = normalise(logw)
    out["cos"] = cos
    if state is not None:
        out["est"] = estimate(out["w"], state)
    return out


def exp_f32(x: np.ndarray) -> np.ndarray:
    """fp32 -> fp32 exponential evaluated in fp64 and rounded once (D14): independent of the
    libm that computes it, so CPU and GPU produce the same bits."""
    return np.exp(np.asarray(x, dtype=np.float32).astype(np.float64)).astype(np.float32)


def normalise(logw: np.ndarray) -> dict:
    """m = max l; S = sum_fp64 snap36(exp(l - m)); lse = fp32(m + log S); w = exp(l - lse);
    both exponentials are `exp_f32` of an fp32 difference (D14)."""
    logw = logw.astype(np.float32)
    m = np.float32(logw.max())
    e = exp_f32((logw - m).astype(np.float32))
    snapped = np.rint(e.astype(np.float64) * 2.0 ** SUM_GRID_BITS) * 2.0 ** -SUM_GRID_BITS
    ssum = float(snapped.sum())
    lse = np.float32(np.float64(m) + np.log(ssum))
    w = exp_f32((logw - lse).astype(np.float32))
    ess = float(1.0 / np.sum(w.astype(np.float64) ** 2))
    return {"logw": logw, "w": w, "lse": lse, "max": m, "sumexp": ssum, "ess": ess}


def estimate(w: np.ndarray, state: np.ndarray) -> np.ndarray:
    """Weighted mean of (cx, cy, s) *before* resampling, fp64 accumulate -> fp32 [3] (D10)."""
    wd = w.astype(np.float64)
    tot = wd.sum()
    return np.array([(wd * state[r].astype(np.float64)).sum() / tot for r in (CX, CY, S)],
                    dtype=np.float32)


def snap_weights(w: np.ndarray) -> np.ndarray:
    """q_i = rint(float64(w_i) * 2^40) * 2^-40 (round-half-even), fp64 (D6)."""
    return np.rint(w.astype(np.float64) * 2.0 ** GRID_BITS) * 2.0 ** -GRID_BITS


def systematic_resample(w: np.ndarray, u: float) -> np.ndarray:
    """Systematic resampling with one uniform draw (D6).  Returns int32 ancestor indices [N].

    C = inclusive cumsum(q) in fp64 (exact, see GRID_BITS); step = C[N-1] / N;
    p_j = (j + u) * step; a_j = #{i : C_i <= p_j} clamped to N-1.
    """
    n = w.shape[0]
    c = np.cumsum(snap_weights(w))
    step = c[-1] / np.float64(n)
    p = (np.arange(n, dtype=np.float64) + np.float64(u)) * step
    a = np.searchsorted(c, p, side="right")
    return np.minimum(a, n - 1).astype(np.int32)


def gather_state(state: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """state' = state[:, idx] (weights are implicitly reset to 1/N) (D8 of the step order, a8)."""
    return np.ascontiguousarray(state[:, idx]).astype(np.float32)
