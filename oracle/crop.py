"""Per-particle crop + bilinear resize + normalise (oracle, numpy/torch CPU) -- test infra only.

Restates SURVEY.md Appendix A D3/D4 (anchor: /root/reference/README.md:7 -- the input side of
"Vision Transformer ... feature extraction"; the reference ships no code).

Output pixel (i, j), i, j in [0, S), S = crop size (224 by default):
    xs = cx - w/2 + (j + 0.5) * w/S - 0.5          ys = cy - h/2 + (i + 0.5) * h/S - 0.5
clamped to [0, W-1] x [0, H-1] (border replicate), bilinear between floor / floor+1 neighbours,
no antialias; then ``(v/255 - 0.5)/0.5``.  All arithmetic in fp32.

``size`` (a multiple of 16, 16..224) is the SURVEY.md section 8(f)-2 "native-resolution crops"
revision of the oracle: the same sampling rule on an S x S grid, giving (S/16)^2 patches.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CROP = 224
PATCH = 16
GRID = CROP // PATCH          # 14
N_PATCH = GRID * GRID         # 196
PATCH_K = 3 * PATCH * PATCH   # 768


def boxes_from_state(state: np.ndarray, w0: float, h0: float) -> np.ndarray:
    """[6,N] state -> [N,4] boxes (cx, cy, w, h) with w = w0*exp(s), h = h0*exp(s) (D5)."""
    # D14: exp evaluated in fp64, rounded once to fp32 (libm-independent, bit-equal on the GPU)
    es = np.exp(state[4].astype(np.float32).astype(np.float64)).astype(np.float32)
    return np.stack([state[0], state[1], np.float32(w0) * es, np.float32(h0) * es],
                    axis=1).astype(np.float32)


def check_crop_size(size: int) -> int:
    size = int(size)
    if size < PATCH or size > CROP or size % PATCH:
        raise ValueError(f"crop size must be a multiple of {PATCH} in [{PATCH}, {CROP}], got {size}")
    return size


def _axis_taps(c, size, extent, out=CROP):
    """fp32 sample coordinates along one axis -> (i0, i1, frac)."""
    k = (np.arange(out, dtype=np.float32) + np.float32(0.5))
    step = (size / np.float32(out)).astype(np.float32)                     # [N]
    base = (c - np.float32(0.5) * size - np.float32(0.5)).astype(np.float32)
    xs = (base[:, None] + k[None, :] * step[:, None]).astype(np.float32)   # [N,S]
    xs = np.clip(xs, np.float32(0.0), np.float32(extent - 1))
    x0 = np.floor(xs)
    fr = (xs - x0).astype(np.float32)
    i0 = x0.astype(np.int64)
    i1 = np.minimum(i0 + 1, extent - 1)
    return i0, i1, fr


def crop_resize(frame: np.ndarray, boxes: np.ndarray, chunk: int = 64,
                size: int = CROP) -> np.ndarray:
    """frame uint8 [H,W,3], boxes fp32 [N,4] (cx,cy,w,h) -> fp32 [N,3,S,S] normalised."""
    S = check_crop_size(size)
    frame = np.asarray(frame)
    assert frame.dtype == np.uint8 and frame.ndim == 3 and frame.shape[2] == 3
    H, W, _ = frame.shape
    boxes = np.asarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    out = np.empty((n, 3, S, S), dtype=np.float32)
    img = frame.astype(np.float32)
    for s in range(0, n, chunk):
        b = boxes[s:s + chunk]
        x0, x1, fx = _axis_taps(b[:, 0], b[:, 2], W, S)
        y0, y1, fy = _axis_taps(b[:, 1], b[:, 3], H, S)
        # gather the four neighbours: [n, S(i), S(j), 3]
        p00 = img[y0[:, :, None], x0[:, None, :]]
        p01 = img[y0[:, :, None], x1[:, None, :]]
        p10 = img[y1[:, :, None], x0[:, None, :]]
        p11 = img[y1[:, :, None], x1[:, None, :]]
        fxx = fx[:, None, :, None]
        fyy = fy[:, :, None, None]
        one = np.float32(1.0)
        top = (one - fxx) * p00 + fxx * p01
        bot = (one - fxx) * p10 + fxx * p11
        v = (one - fyy) * top + fyy * bot
        v = v * np.float32(2.0 / 255.0) - one
        out[s:s + chunk] = np.transpose(v, (0, 3, 1, 2))
    return out


def crop_resize_grid_sample(frame: np.ndarray, boxes: np.ndarray, size: int = CROP) -> np.ndarray:
    """Independent formulation through ``F.grid_sample`` (cross-check of :func:`crop_resize`)."""
    H, W, _ = frame.shape
    S = check_crop_size(size)
    img = torch.from_numpy(np.ascontiguousarray(frame)).permute(2, 0, 1).float()[None]
    boxes_t = torch.from_numpy(np.asarray(boxes, dtype=np.float32)).double()
    k = torch.arange(S, dtype=torch.float64) + 0.5
    xs = boxes_t[:, 0:1] - boxes_t[:, 2:3] / 2 + k[None] * boxes_t[:, 2:3] / S - 0.5
    ys = boxes_t[:, 1:2] - boxes_t[:, 3:4] / 2 + k[None] * boxes_t[:, 3:4] / S - 0.5
    gx = (2 * xs + 1) / W - 1
    gy = (2 * ys + 1) / H - 1
    n = boxes_t.shape[0]
    grid = torch.stack([gx[:, None, :].expand(n, S, S),
                        gy[:, :, None].expand(n, S, S)], dim=-1)
    out = F.grid_sample(img.double().expand(n, -1, -1, -1), grid, mode="bilinear",
                        padding_mode="border", align_corners=False)
    out = (out / 255.0 - 0.5) / 0.5
    return out.float().numpy()


def patchify(crops: np.ndarray) -> np.ndarray:
    """[N,3,S,S] -> patch-major GEMM-A layout [N*(S/16)^2, 768], k = c*256 + ky*16 + kx."""
    n, _, S, _ = crops.shape
    g = S // PATCH
    x = crops.reshape(n, 3, g, PATCH, g, PATCH)
    x = np.transpose(x, (0, 2, 4, 1, 3, 5))
    return np.ascontiguousarray(x).reshape(n * g * g, PATCH_K)
