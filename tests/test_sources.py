"""Frame sources (vitparticlefiltertracker_b200/sources.py): host-only input adapters."""
import os

import numpy as np
import pytest

from vitparticlefiltertracker_b200.sources import array_frames, image_frames, open_source, video_frames


def _clip(n=4, H=24, W=32, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)


def test_array_and_npy_sources(tmp_path):
    clip = _clip()
    assert all(np.array_equal(a, b) for a, b in zip(array_frames(clip), clip))
    p = tmp_path / "clip.npy"
    np.save(p, clip)
    got = list(open_source("npy", str(p)))
    assert len(got) == 4 and all(g.flags["C_CONTIGUOUS"] and np.array_equal(g, c) for g, c in zip(got, clip))
    with pytest.raises(ValueError):
        list(array_frames(np.zeros((2, 8, 8), np.uint8)))


def test_image_sequence_is_decoded_as_rgb_in_sorted_order(tmp_path):
    cv2 = pytest.importorskip("cv2")
    clip = _clip(3)
    for i, fr in enumerate(clip):
        assert cv2.imwrite(str(tmp_path / f"f{i:03d}.png"), fr[..., ::-1])      # OpenCV writes BGR
    got = list(image_frames(str(tmp_path / "f*.png")))
    assert len(got) == 3 and all(np.array_equal(g, c) for g, c in zip(got, clip))   # PNG is lossless
    with pytest.raises(FileNotFoundError):
        list(image_frames(str(tmp_path / "none*.png")))
    cv2.imwrite(str(tmp_path / "f999.png"), np.zeros((5, 5, 3), np.uint8))
    with pytest.raises(ValueError):
        list(image_frames(str(tmp_path / "f*.png")))


def test_video_source_errors_and_unknown_kind(tmp_path):
    pytest.importorskip("cv2")
    with pytest.raises(FileNotFoundError):
        next(video_frames(str(tmp_path / "missing.mp4")))
    with pytest.raises(ValueError):
        open_source("carrier-pigeon", "x")


def test_video_file_round_trip_if_a_codec_is_available(tmp_path):
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (32, 24))
    if not w.isOpened():
        pytest.skip("no MJPG encoder in this OpenCV build")
    clip = np.repeat(np.repeat(_clip(5, 6, 8), 4, axis=1), 4, axis=2)            # blocky: survives JPEG
    for fr in clip:
        w.write(np.ascontiguousarray(fr[..., ::-1]))
    w.release()
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        pytest.skip("video file was not written")
    got = list(video_frames(path, max_frames=4))
    assert len(got) == 4 and got[0].shape == (24, 32, 3) and got[0].dtype == np.uint8
    assert np.abs(got[0].astype(int) - clip[0].astype(int)).mean() < 12           # lossy codec
