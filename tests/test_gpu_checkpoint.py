"""GPU: the real-checkpoint path end to end.

A timm-keyed checkpoint with everything a real one has (tests/checkpoint_util.py: wrapper,
`module.` prefix, head / fc_norm extras, fp16 tensors, non-trivial biases and LayerNorm
parameters, outlier channels) goes through `Tracker(state_dict=...)` and through `main.py`'s
`checkpoint:` key, and must match `OracleTracker(state_dict=...)` within the stated tolerance."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

from checkpoint_util import realistic_checkpoint
from gpu_util import cuda
from oracle import OracleTracker, pf
from parity_cases import TOL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model,n", [("vit_tiny_patch16_224", 128), ("vit_base_patch16_224", 32)])
def test_tracker_with_checkpoint_matches_oracle_with_the_same_checkpoint(model, n, tmp_path):
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    ckpt, _ = realistic_checkpoint(model, seed=5)
    path = str(tmp_path / "ckpt.pt")
    torch.save(ckpt, path)
    frames, gt = make_clip("box", 224, 224, 3, seed=0, target=(48, 48))
    ora = OracleTracker(model, n, seed=0, state_dict=ckpt)
    trk = Tracker(model, n, seed=0, state_dict=path)              # by path, as main.py passes it
    seeded = Tracker(model, n, seed=0)                            # the default seeded init
    for t in (ora, trk, seeded):
        t.init(frames[0], gt[0])
    assert np.array_equal(trk.particles.cpu().numpy(), ora.particles)
    cos_t = torch.nn.functional.cosine_similarity(trk.template.cpu(), torch.from_numpy(ora.template), 0)
    assert cos_t > 1 - 1e-3
    # the checkpoint is really in use: its template differs from the seeded-init tracker's
    assert torch.nn.functional.cosine_similarity(trk.template, seeded.template, 0) < 0.999
    tol = TOL["bf16"]
    for t in (1, 2):
        trk.particles.copy_(torch.from_numpy(ora.particles))      # teacher forcing
        so, sg = ora.step(frames[t]), trk.step(frames[t])
        dcos = np.abs(trk.logw.cpu().numpy() * np.float32(0.1) - ora.last["cos"]).max()
        wg = trk.weights.cpu().numpy()
        relw = np.abs(wg / ora.last["w"] - 1).max()
        print(f"{model} frame {t}: |dcos| {dcos:.2e}, weights rel {relw:.2e}")
        assert dcos <= tol["dcos"] and relw <= tol["relw"]
        assert np.abs(np.array(sg.bbox) - np.array(so.bbox)).max() <= tol["bbox"]
        assert abs(sg.ess / so.ess - 1) <= tol["ess"]
        assert np.array_equal(trk.idx.cpu().numpy(), pf.systematic_resample(wg, trk.last_u))


def test_main_py_checkpoint_key_loads_the_file(tmp_path):
    """`checkpoint:` in config.yaml (README.md:42) -> the same track as Tracker(state_dict=...)."""
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    model = "vit_tiny_patch16_224"
    ckpt, _ = realistic_checkpoint(model, seed=9)
    path = str(tmp_path / "ckpt.pt")
    torch.save(ckpt, path)
    syn = dict(clip="box", height=120, width=160, frames=4, target=[32, 24], seed=3)
    cfg = dict(model=model, n_particles=64, seed=1, checkpoint=path,
               input=dict(kind="synthetic", synthetic=syn))
    p = tmp_path / "cfg.yaml"
    p.write_text(yaml.safe_dump(cfg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(p)], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    got = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    frames, gt = make_clip("box", 120, 160, 4, 3, (32, 24))
    trk = Tracker(model, 64, seed=1, state_dict=ckpt)
    trk.init(frames[0], gt[0])
    want = [trk.step(frames[t]) for t in (1, 2, 3)]
    assert [g["frame"] for g in got] == [1, 2, 3]
    for g, w in zip(got, want):
        assert np.allclose(g["bbox"], w.bbox, atol=2e-3)          # main.py rounds to 3 decimals
    # and a checkpoint for the wrong model is refused with the key's name
    cfg["model"] = "vit_small_patch16_224"
    p.write_text(yaml.safe_dump(cfg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(p)], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "patch_embed.proj.weight" in r.stderr
