"""GPU: tracking-accuracy regression ("a tracking-accuracy metric (IoU /
centre error vs synthetic ground truth)"; upstream anchors README.md:3 "precise", :9 "high
accuracy", :48 needle tracking).  Free-running filters on the box and the needle clip, GPU and
CPU oracle side by side: both must hold the target, and the GPU track must stay with the
oracle's (the encoder's 16-bit rounding perturbs weights by ~1e-2, which free-running
resampling turns into slightly different, statistically equivalent particle sets)."""
import numpy as np
import pytest

from gpu_util import cuda
from oracle import OracleTracker

pytestmark = pytest.mark.gpu

#: (clip kind, H, W, target, velocity, min mean IoU, max mean centre error in pixels)
CLIPS = {
    "box": ("box", 224, 224, (48, 48), (1.5, 0.5), 0.60, 10.0),
    "needle": ("needle", 240, 320, (3, 120), (1.0, 0.5), 0.40, 24.0),
}


def _score(trk, frames, gt):
    from vitparticlefiltertracker_b200.synthetic import iou
    trk.init(frames[0], gt[0])
    ious, cerr, boxes = [], [], []
    for t in range(1, len(frames)):
        s = trk.step(frames[t])
        ious.append(iou(s.bbox, gt[t]))
        cerr.append(float(np.hypot(s.bbox[0] + s.bbox[2] / 2 - gt[t][0] - gt[t][2] / 2,
                                   s.bbox[1] + s.bbox[3] / 2 - gt[t][1] - gt[t][3] / 2)))
        boxes.append(s.bbox)
    return float(np.mean(ious)), float(np.mean(cerr)), np.array(boxes)


@pytest.mark.parametrize("clip", list(CLIPS))
def test_tracking_accuracy_gpu_and_oracle_side_by_side(clip):
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    kind, H, W, target, vel, min_iou, max_err = CLIPS[clip]
    frames, gt = make_clip(kind, H, W, 11, seed=0, target=target, velocity=vel)
    n, model = 128, "vit_tiny_patch16_224"
    iou_o, err_o, box_o = _score(OracleTracker(model, n, seed=0), frames, gt)
    iou_g, err_g, box_g = _score(Tracker(model, n, seed=0), frames, gt)
    drift = float(np.abs(box_g - box_o).max())
    print(f"{clip}: mean IoU gpu {iou_g:.3f} / oracle {iou_o:.3f}; centre error gpu {err_g:.2f} / "
          f"oracle {err_o:.2f} px; max |bbox_gpu - bbox_oracle| {drift:.2f} px")
    assert iou_o >= min_iou and err_o <= max_err          # the oracle holds the target ...
    assert iou_g >= min_iou and err_g <= max_err          # ... and so does the GPU path
    # free-running tracks are two realisations of the same filter (1e-2 weight perturbations change
    # which particles are duplicated): they must agree statistically, not sample by sample
    assert abs(iou_g - iou_o) <= 0.08 and abs(err_g - err_o) <= 3.0
    assert drift <= 8.0
