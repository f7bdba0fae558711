"""Generates tests/golden/oracle_v2_crop.npz: golden vectors of the oracle's variable crop size
revision ("native-resolution crops"): crop_resize / patchify at S = 64 and
S = 112, the resampled position embedding, the ViT embedding of 17- and 50-token inputs, and a
short OracleTracker(crop_size=64) trajectory.  The 224-pixel vectors stay in oracle_v1.npz.

Run from the repo root:  python tests/golden/make_golden_crop.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.crop import crop_resize, patchify  # noqa: E402
from oracle.tracker import OracleTracker  # noqa: E402
from oracle.vit import init_vit_weights, resample_pos_embed, vit_forward  # noqa: E402
from vitparticlefiltertracker_b200.synthetic import make_clip  # noqa: E402

MODEL = "vit_tiny_patch16_224"


def main():
    out = {}
    rng = np.random.default_rng(43)
    frame = rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
    boxes = np.array([[64, 48, 40, 30], [5, 90, 33.5, 61.25], [100.3, 20.7, 200, 150],
                      [60, 40, 64, 64]], np.float32)
    out["crop_frame"] = frame
    out["crop_boxes"] = boxes
    sd = init_vit_weights(MODEL)
    for S in (64, 112):
        crops = crop_resize(frame, boxes, size=S)
        out[f"crop_samples_{S}"] = crops[:, :, ::5, ::7].copy()
        out[f"crop_checksum_{S}"] = np.array([np.float64(crops.astype(np.float64).sum()),
                                              np.float64((crops.astype(np.float64) ** 2).sum())])
        out[f"patch_samples_{S}"] = patchify(crops)[::3, ::41].copy()
        out[f"pos_embed_{S}"] = resample_pos_embed(sd["pos_embed"], S // 16)[0, :, :8].numpy()
        out[f"vit_emb_{S}"] = vit_forward(crops, sd, MODEL)
    frames, gt = make_clip("box", 224, 224, 4, seed=0, target=(48, 48))
    trk = OracleTracker(MODEL, 16, seed=0, crop_size=64)
    trk.init(frames[0], gt[0])
    traj = []
    for t in range(1, 4):
        s = trk.step(frames[t])
        traj.append([*s.bbox, s.ess])
    out["track_traj_64"] = np.array(traj, dtype=np.float64)
    out["track_particles_64"] = trk.particles
    out["track_idx_64"] = trk.last["idx"]
    out["track_template_64"] = trk.template
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v2_crop.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
