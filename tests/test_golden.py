"""The oracle against its committed golden vectors (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import pf
from oracle.crop import crop_resize, patchify
from oracle.philox import philox4x32_10, philox_normals, philox_uniform53
from oracle.tracker import OracleTracker
from oracle.vit import init_vit_weights, vit_forward


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "oracle_v1.npz"))


def test_philox_golden(G):
    ids = np.arange(5, 13, dtype=np.uint32)
    assert np.array_equal(np.stack(philox4x32_10((ids, 3, 1, 0), (0xDEADBEEF, 0x12345678))),
                          G["philox_words"])
    assert np.allclose(philox_normals(np.arange(100, 116), 7, 2, 99), G["normals"], atol=1e-6)
    assert np.array_equal(np.array([philox_uniform53(t, 0, 0) for t in range(1, 9)]), G["uniforms"])


def test_crop_and_vit_golden(G):
    crops = crop_resize(G["crop_frame"], G["crop_boxes"])
    assert np.allclose(crops[:, :, ::13, ::17], G["crop_samples"], atol=1e-6)
    cs = np.array([crops.astype(np.float64).sum(), (crops.astype(np.float64) ** 2).sum()])
    assert np.allclose(cs, G["crop_checksum"], rtol=1e-9)
    assert np.allclose(patchify(crops)[::37, ::41], G["patch_samples"], atol=1e-6)
    sd = init_vit_weights("vit_tiny_patch16_224")
    probe = np.array([sd["blocks.0.attn.qkv.weight"][0, :8].numpy(),
                      sd["blocks.11.mlp.fc2.weight"][5, :8].numpy(),
                      sd["pos_embed"][0, 100, :8].numpy()])
    assert np.array_equal(probe, G["weight_probe"])
    emb = vit_forward(crops, sd, "vit_tiny_patch16_224")
    assert np.allclose(emb, G["vit_emb"], atol=2e-4)      # CPU BLAS summation-order slack


def test_pf_golden(G):
    nm = pf.normalise(G["norm_logw"])
    assert np.allclose(nm["w"], G["norm_w"], rtol=1e-6)
    assert np.allclose([nm["lse"], nm["ess"]], G["norm_lse_ess"], rtol=1e-6)
    got = np.stack([pf.systematic_resample(G["norm_w"], u) for u in (0.0, 0.37, 0.999)])
    assert np.array_equal(got, G["resample_idx"])


def test_tracker_golden(G):
    from vitparticlefiltertracker_b200.synthetic import make_clip
    frames, gt = make_clip("box", 224, 224, 4, seed=0, target=(48, 48))
    assert int(frames.astype(np.int64).sum()) == int(G["clip_checksum"][0])
    assert np.array_equal(gt, G["clip_gt"])
    trk = OracleTracker("vit_tiny_patch16_224", 16, seed=0)
    trk.init(frames[0], gt[0])
    assert np.allclose(trk.template, G["track_template"], atol=2e-4)
    traj = []
    for t in range(1, 4):
        s = trk.step(frames[t])
        traj.append([*s.bbox, s.ess])
    assert np.allclose(np.array(traj), G["track_traj"], rtol=2e-3, atol=2e-3)
