"""Synthetic moving-target clips (measurement fixture).

The reference names its inputs only as "video feed or images" (/root/reference/README.md:42) and
its motivating case as needle tracking (/root/reference/README.md:48); it ships no data.  There is
no network here either, so benchmarks and tests track a seeded synthetic target with known
ground truth.  Host-side numpy; not part of the timed region.
"""
from __future__ import annotations

import math

import numpy as np


def _value_noise(rng, H, W, cell=8):
    """Smooth textured background: coarse random grid, bilinear upsample, plus fine grain."""
    gh, gw = H // cell + 2, W // cell + 2
    coarse = rng.integers(32, 224, size=(gh, gw, 3)).astype(np.float32)
    ys = (np.arange(H, dtype=np.float32) + 0.5) / cell
    xs = (np.arange(W, dtype=np.float32) + 0.5) / cell
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    a = coarse[y0][:, x0]
    b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]
    d = coarse[y0 + 1][:, x0 + 1]
    img = (1 - fy) * ((1 - fx) * a + fx * b) + fy * ((1 - fx) * c + fx * d)
    img += rng.integers(-12, 13, size=(H, W, 3)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_clip(kind="box", H=224, W=224, n_frames=32, seed=0, target=(48, 48), velocity=None):
    """Returns (frames uint8 [F,H,W,3], gt float32 [F,4] as (x, y, w, h) top-left boxes).

    kind="box":    random-texture rectangle, centre c(t) = c0 + v t + A sin(2 pi t / P).
    kind="needle": thin bright bar (3 px x 120 px) rotating 1 degree per frame about a moving
                   centre; gt is the axis-aligned bounding box of the bar.
    velocity: optional (vx, vy) px/frame of the linear part (default: cross 30 % of the width
              and 10 % of the height over the clip).
    """
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    bg = _value_noise(rng, H, W)
    tw, th = int(target[0]), int(target[1])
    frames = np.empty((n_frames, H, W, 3), dtype=np.uint8)
    gt = np.empty((n_frames, 4), dtype=np.float32)
    c0 = np.array([W * 0.35, H * 0.45], dtype=np.float64)
    margin = max(tw, th)
    span = np.array([W * 0.3, H * 0.1])
    vel = span / max(n_frames - 1, 1)
    if velocity is not None:                      # explicit (vx, vy) in pixels per frame
        vel = np.array(velocity, dtype=np.float64)
    amp = np.array([0.0, min(H * 0.08, margin * 0.5)])
    period = 24.0
    if kind == "box":
        tex = rng.integers(0, 256, size=(th, tw, 3)).astype(np.uint8)
        # keep some low-frequency structure so that bilinear upsampling preserves identity
        tex[: th // 2, : tw // 2] = (tex[: th // 2, : tw // 2] // 4 + 192).astype(np.uint8)
        tex[th // 2:, tw // 2:] = (tex[th // 2:, tw // 2:] // 4).astype(np.uint8)
    for t in range(n_frames):
        c = c0 + vel * t + amp * math.sin(2 * math.pi * t / period)
        f = bg.copy()
        if kind == "box":
            x = int(round(c[0] - tw / 2))
            y = int(round(c[1] - th / 2))
            x = min(max(x, 0), W - tw)
            y = min(max(y, 0), H - th)
            f[y:y + th, x:x + tw] = tex
            gt[t] = (x, y, tw, th)
        elif kind == "needle":
            L, T = 120.0, 3.0
            ang = math.radians(20.0 + t)
            dx, dy = math.cos(ang), math.sin(ang)
            r = int(L / 2 + 3)
            yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
            along = xx * dx + yy * dy
            across = -xx * dy + yy * dx
            mask = (np.abs(along) <= L / 2) & (np.abs(across) <= T / 2)
            cx, cy = int(round(c[0])), int(round(c[1]))
            ys = np.clip(cy + yy[mask], 0, H - 1)
            xs = np.clip(cx + xx[mask], 0, W - 1)
            f[ys, xs] = np.array([240, 240, 250], dtype=np.uint8)
            bw = abs(dx) * L + abs(dy) * T
            bh = abs(dy) * L + abs(dx) * T
            gt[t] = (cx - bw / 2, cy - bh / 2, bw, bh)
        else:
            raise ValueError(f"unknown clip kind {kind!r}")
        frames[t] = f
    return frames, gt


def iou(a, b) -> float:
    """IoU of two (x, y, w, h) boxes."""
    ax0, ay0, ax1, ay1 = a[0], a[1], a[0] + a[2], a[1] + a[3]
    bx0, by0, bx1, by1 = b[0], b[1], b[0] + b[2], b[1] + b[3]
    iw = max(0.0, min(ax1, bx1) - max(ax0, bx0))
    ih = max(0.0, min(ay1, by1) - max(ay0, by0))
    inter = iw * ih
    union = a[2] * a[3] + b[2] * b[3] - inter
    return float(inter / union) if union > 0 else 0.0
