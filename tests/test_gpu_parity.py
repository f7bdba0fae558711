"""GPU parity at the tracker level, with the stated tolerances ASSERTED:

* teacher-forced steps against the fp32 oracle for ViT-Ti / ViT-B / ViT-L and the configs[1]
  shape, bf16 (default) and fp16 operands, fp32 and fp16 residual stream;
* a free-running 8-frame run in which the oracle is fed the GPU's embeddings: everything that
  is not the encoder -- motion model, likelihood, normalisation, CDF, resampling, gather -- must
  then agree BIT FOR BIT, frame after frame (D14).
"""
import numpy as np
import pytest
import torch

from gpu_util import cuda
from parity_cases import CASES, TOL, gpu_metrics

pytestmark = pytest.mark.gpu

DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16}


@pytest.mark.parametrize("op", ["bf16", "fp16"])
@pytest.mark.parametrize("case", list(CASES))
def test_teacher_forced_step_within_stated_tolerance(case, op):
    cuda()
    m = gpu_metrics(case, dtype=DTYPES[op])
    print(m)
    tol = TOL[op]
    assert m["init_exact"] and m["propagate_exact"] and m["resample_exact"], m
    assert m["crop_exact"], m
    assert m["template_cos"] > 1 - 1e-3
    for k in ("dcos", "relw", "ess", "bbox", "emb"):
        assert m[k] <= tol[k], (k, m[k], tol[k])


@pytest.mark.parametrize("case", ["vit_tiny_256", "vit_base_64"])
def test_fp16_residual_stream_stays_inside_the_bf16_tolerance(case):
    """The fp16 residual stream (default since round 2: half the residual bytes) adds ~2 % to the
    bf16 operands' own distance from the oracle; the stated bf16 tolerance holds with it."""
    cuda()
    a = gpu_metrics(case, residual_dtype=torch.float32)
    b = gpu_metrics(case, residual_dtype=torch.float16)
    print(a, b)
    for k in ("dcos", "relw", "ess", "bbox", "emb"):
        assert b[k] <= TOL["bf16"][k], (k, b[k])
    assert b["emb"] < 1.25 * a["emb"] + 1e-4
    assert b["crop_exact"] and b["propagate_exact"] and b["resample_exact"]


@pytest.mark.parametrize("graph", [False, True])
def test_free_running_filter_is_bit_exact_given_the_gpu_embeddings(graph):
    """8 free-running frames; the oracle gets the GPU's embeddings (`_emb=`) and nothing else.
    Particles, log-weights, weights, ancestor indices: bit-identical every frame."""
    from oracle import OracleTracker
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    frames, gt = make_clip("box", 224, 224, 9, seed=0, target=(48, 48))
    n = 256
    ora = OracleTracker("vit_tiny_patch16_224", n, seed=0)
    trk = Tracker("vit_tiny_patch16_224", n, seed=0, cuda_graph=graph)
    ora.init(frames[0], gt[0])
    trk.init(frames[0], gt[0])
    assert np.array_equal(trk.particles.cpu().numpy(), ora.particles)
    ora.template = trk.template.cpu().numpy()          # the encoder's output is the only input
    for t in range(1, 9):
        sg = trk.step(frames[t])
        so = ora.step(frames[t], _emb=trk.emb.cpu().numpy())
        assert np.array_equal(trk.logw.cpu().numpy(), ora.last["logw"]), t
        assert np.array_equal(trk.weights.cpu().numpy(), ora.last["w"]), t
        assert np.array_equal(trk.idx.cpu().numpy(), ora.last["idx"]), t
        assert np.array_equal(trk.particles.cpu().numpy(), ora.particles), t
        assert trk.last_u == ora.last["u"]
        assert np.abs(np.array(sg.bbox) - np.array(so.bbox)).max() < 1e-4     # oracle est is fp32
        assert abs(sg.ess / so.ess - 1) < 1e-9
