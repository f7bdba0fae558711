"""Generates tests/golden/oracle_v1.npz: small fixed input/output vectors of the CPU oracle.

The upstream repository ships no golden vectors (it is a README only), so these pin the oracle
itself: any later edit that moves its results fails tests/test_golden.py.  Run from the repo
root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pf  # noqa: E402
from oracle.crop import crop_resize, patchify  # noqa: E402
from oracle.philox import philox4x32_10, philox_normals, philox_uniform53  # noqa: E402
from oracle.tracker import OracleTracker  # noqa: E402
from oracle.vit import init_vit_weights, vit_forward  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402


def main():
    out = {}
    ids = np.arange(5, 13, dtype=np.uint32)
    out["philox_words"] = np.stack(philox4x32_10((ids, 3, 1, 0), (0xDEADBEEF, 0x12345678)))
    out["normals"] = philox_normals(np.arange(100, 116), 7, 2, 99)
    out["uniforms"] = np.array([philox_uniform53(t, 0, 0) for t in range(1, 9)])
    rng = np.random.default_rng(42)
    frame = rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
    boxes = np.array([[64, 48, 40, 30], [5, 90, 33.5, 61.25], [100.3, 20.7, 200, 150]], np.float32)
    crops = crop_resize(frame, boxes)
    out["crop_frame"] = frame
    out["crop_boxes"] = boxes
    out["crop_samples"] = crops[:, :, ::13, ::17].copy()          # subsampled [3,3,18,14]
    out["crop_checksum"] = np.array([np.float64(crops.astype(np.float64).sum()),
                                     np.float64((crops.astype(np.float64) ** 2).sum())])
    out["patch_samples"] = patchify(crops)[::37, ::41].copy()
    sd = init_vit_weights("vit_tiny_patch16_224")
    out["weight_probe"] = np.array([sd["blocks.0.attn.qkv.weight"][0, :8].numpy(),
                                    sd["blocks.11.mlp.fc2.weight"][5, :8].numpy(),
                                    sd["pos_embed"][0, 100, :8].numpy()])
    out["vit_emb"] = vit_forward(crops, sd, "vit_tiny_patch16_224")
    lw = (rng.standard_normal(512) * 5).astype(np.float32)
    nm = pf.normalise(lw)
    out["norm_logw"] = lw
    out["norm_w"] = nm["w"]
    out["norm_lse_ess"] = np.array([nm["lse"], nm["ess"]], dtype=np.float64)
    out["resample_idx"] = np.stack([pf.systematic_resample(nm["w"], u) for u in (0.0, 0.37, 0.999)])
    frames, gt = make_clip("box", 224, 224, 4, seed=0, target=(48, 48))
    trk = OracleTracker("vit_tiny_patch16_224", 16, seed=0)
    trk.init(frames[0], gt[0])
    traj = []
    for t in range(1, 4):
        s = trk.step(frames[t])
        traj.append([*s.bbox, s.ess])
    out["clip_checksum"] = np.array([int(frames.astype(np.int64).sum())])
    out["clip_gt"] = gt
    out["track_traj"] = np.array(traj, dtype=np.float64)
    out["track_particles"] = trk.particles
    out["track_idx"] = trk.last["idx"]
    out["track_template"] = trk.template
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
