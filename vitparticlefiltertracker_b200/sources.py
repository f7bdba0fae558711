"""Frame sources for the tracker's input side (upstream anchor: README.md:42, "provide the input
data (e.g., video feed or images)").  Host-only helpers that yield uint8 [H, W, 3] **RGB** numpy
frames; they sit before the hot path and use OpenCV only when a video
file / camera or an encoded image actually has to be decoded."""
from __future__ import annotations

import glob
import os

import numpy as np


def _cv2():
    try:
        import cv2
    except ImportError as e:                       # pragma: no cover
        raise RuntimeError("video / image input needs OpenCV (cv2)") from e
    return cv2


def array_frames(frames):
    """uint8 [F, H, W, 3] array (or memory-mapped .npy) -> frames."""
    for f in range(len(frames)):
        fr = np.ascontiguousarray(frames[f])
        if fr.dtype != np.uint8 or fr.ndim != 3 or fr.shape[2] != 3:
            raise ValueError("frames must be uint8 [F, H, W, 3] RGB")
        yield fr


def image_frames(pattern: str):
    """Sorted glob of encoded images (PNG / JPEG / ...) -> RGB frames of one common size."""
    cv2 = _cv2()
    paths = sorted(glob.glob(pattern))
    if not paths:
        raise FileNotFoundError(f"no image matches {pattern!r}")
    shape = None
    for p in paths:
        bgr = cv2.imread(p, cv2.IMREAD_COLOR)
        if bgr is None:
            raise ValueError(f"cannot decode {p}")
        if shape is None:
            shape = bgr.shape
        elif bgr.shape != shape:
            raise ValueError(f"{p}: frame size {bgr.shape[:2]} differs from the first frame's {shape[:2]}")
        yield np.ascontiguousarray(bgr[..., ::-1])


def video_frames(source, max_frames: int | None = None):
    """A video file path or a camera index -> RGB frames (cv2.VideoCapture)."""
    cv2 = _cv2()
    if isinstance(source, str) and not source.isdigit() and not os.path.exists(source):
        raise FileNotFoundError(source)
    cap = cv2.VideoCapture(int(source) if isinstance(source, str) and source.isdigit() else source)
    if not cap.isOpened():
        raise RuntimeError(f"cannot open video source {source!r}")
    n = 0
    try:
        while max_frames is None or n < max_frames:
            ok, bgr = cap.read()
            if not ok:
                break
            n += 1
            yield np.ascontiguousarray(bgr[..., ::-1])
    finally:
        cap.release()


def open_source(kind: str, path=None, **kw):
    """kind: 'npy' (uint8 [F,H,W,3] file), 'images' (glob pattern), 'video' (file path / camera index)."""
    if kind == "npy":
        return array_frames(np.load(path, mmap_mode="r"))
    if kind == "images":
        return image_frames(path)
    if kind == "video":
        return video_frames(path, kw.get("max_frames"))
    raise ValueError(f"unknown input kind {kind!r}: choose npy, images or video")
