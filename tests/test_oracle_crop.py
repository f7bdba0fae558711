import numpy as np

from oracle.crop import (crop_resize, crop_resize_grid_sample, patchify, boxes_from_state, CROP)


def _frame(H=96, W=128, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(H, W, 3), dtype=np.uint8)


def test_matches_grid_sample_formulation():
    f = _frame()
    boxes = np.array([[64, 48, 40, 30], [10, 10, 50, 50], [120, 90, 64, 64], [64.3, 47.7, 224, 224],
                      [30.25, 60.5, 17.3, 9.1]], dtype=np.float32)
    a = crop_resize(f, boxes)
    b = crop_resize_grid_sample(f, boxes)
    assert a.shape == (5, 3, CROP, CROP) and a.dtype == np.float32
    assert np.abs(a - b).max() < 2e-3


def test_identity_box_reproduces_frame():
    f = _frame(224, 224, 1)
    out = crop_resize(f, np.array([[112, 112, 224, 224]], dtype=np.float32))[0]
    want = np.transpose(f.astype(np.float32), (2, 0, 1)) * np.float32(2 / 255) - 1
    assert np.abs(out - want).max() < 1e-5


def test_box_outside_frame_is_border_replicate():
    f = _frame(64, 64, 2)
    out = crop_resize(f, np.array([[-500, -500, 20, 20]], dtype=np.float32))[0]
    corner = f[0, 0].astype(np.float32) * np.float32(2 / 255) - 1
    assert np.allclose(out, corner[:, None, None], atol=1e-6)
    out = crop_resize(f, np.array([[900, 32, 20, 20]], dtype=np.float32))[0]
    assert np.allclose(out[:, :, 0], out[:, :, -1], atol=1e-6)   # every column = right edge


def test_range_and_patchify_layout():
    f = _frame()
    crops = crop_resize(f, np.array([[64, 48, 60, 60]], dtype=np.float32))
    assert crops.min() >= -1 - 1e-6 and crops.max() <= 1 + 1e-6
    p = patchify(crops)
    assert p.shape == (196, 768)
    # patch (py=2, px=3), element (c=1, ky=5, kx=7)
    assert p[2 * 14 + 3, 1 * 256 + 5 * 16 + 7] == crops[0, 1, 2 * 16 + 5, 3 * 16 + 7]


def test_boxes_from_state():
    st = np.zeros((6, 2), dtype=np.float32)
    st[0] = [10, 20]; st[1] = [30, 40]; st[4] = [0, np.log(2)]
    b = boxes_from_state(st, 8, 4)
    assert np.allclose(b, [[10, 30, 8, 4], [20, 40, 16, 8]], rtol=1e-6)
