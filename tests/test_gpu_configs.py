"""GPU: the remaining BASELINE configs at reduced / full scale.
config 2/3 shape (ViT-B at full N): size-independent properties; config 4 model (ViT-L): parity;
config 5 (independent needle streams): MultiStreamTracker == S single-stream Trackers."""
import numpy as np
import pytest
import torch

from oracle import pf
from oracle.crop import crop_resize, patchify, boxes_from_state
from oracle.vit import init_vit_weights, vit_forward, VIT_CONFIGS
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def test_vit_large_matches_oracle():
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    model, n = "vit_large_patch16_224", 2
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    boxes = np.array([[160, 120, 64, 64], [100, 80, 90, 40]], dtype=np.float32)
    crops = crop_resize(frame, boxes)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, sd), dev)
    got = enc.forward(torch.from_numpy(patchify(crops)).to(torch.bfloat16).to(dev), n).cpu()
    rel = float((got - want).norm() / want.norm())
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1).min()
    print(f"vit_large: rel-L2 {rel:.3e}, min cos {cos:.6f}")
    assert rel < 1.5e-2 and cos > 1 - 1e-3


def test_full_size_step_properties_vit_base_4096():
    """BASELINE configs[2] at full size: properties that do not need the (317 s/frame) oracle."""
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    dev = cuda()
    frames, gt = make_clip("box", 1080, 1920, 3, seed=0, target=(96, 96))
    n = 4096
    trk = Tracker("vit_base_patch16_224", n, seed=0)
    trk.init(frames[0], gt[0])
    s1 = trk.step(frames[1])
    assert s1.frame_idx == 1 and 1.0 <= s1.ess <= n and np.isfinite(s1.bbox).all()
    w = trk.weights.cpu().numpy()
    assert abs(float(w.astype(np.float64).sum()) - 1.0) < 1e-4 and (w >= 0).all()
    idx = trk.idx.cpu().numpy()
    assert (np.diff(idx) >= 0).all() and idx.min() >= 0 and idx.max() < n       # sortedness
    assert np.array_equal(idx, pf.systematic_resample(w, trk.last_u))           # bit-exact K4
    assert torch.equal(trk.particles, trk.propagated[:, trk.idx.long()])        # gather
    counts = np.bincount(idx, minlength=n)
    q = pf.snap_weights(w)
    assert (np.abs(counts - n * q / q.sum()) < 1 + 1e-6).all()
    # crop parity on a sample of the real particle boxes (K1 at the 1080p frame size)
    st = trk.propagated.cpu().numpy()
    sel = np.arange(0, n, 512)
    want = patchify(crop_resize(frames[1], boxes_from_state(st[:, sel], trk.w0, trk.h0)))
    got = trk.patches.view(n, 196, 768)[torch.from_numpy(sel).to(dev)].reshape(-1, 768).float().cpu().numpy()
    want_bf = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    assert np.abs(got - want_bf).max() <= 2.0 ** -7
    # batch invariance at full size: the first 8 crops alone give bit-identical embeddings
    emb_full = trk.emb[:8].clone()
    small = trk.encoder.forward(trk.patches[:8 * 196].contiguous(), 8)
    assert torch.equal(small, emb_full)
    # embedding parity against the oracle on those 8 crops
    want_emb = vit_forward(crop_resize(frames[1], boxes_from_state(st[:, :8], trk.w0, trk.h0)),
                           init_vit_weights("vit_base_patch16_224"), "vit_base_patch16_224")
    cos = torch.nn.functional.cosine_similarity(emb_full.cpu(), torch.from_numpy(want_emb), dim=1)
    assert cos.min() > 1 - 1e-3
    s2 = trk.step(frames[2])
    assert s2.frame_idx == 2 and np.isfinite(s2.bbox).all()


def test_multistream_equals_independent_trackers():
    from vitparticlefiltertracker_b200 import Tracker, MultiStreamTracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    S, n = 3, 128
    clips = [make_clip("needle", 240, 320, 4, seed=s, target=(3, 120)) for s in range(S)]
    frames = np.stack([c[0] for c in clips], axis=1)            # [F, S, H, W, 3]
    boxes = [c[1][0] for c in clips]
    ms = MultiStreamTracker(S, "vit_tiny_patch16_224", n, seed=4)
    ms.init(frames[0], boxes)
    singles = []
    for s in range(S):
        t = Tracker("vit_tiny_patch16_224", n, seed=4, stream_id=s)
        t.init(frames[0, s], boxes[s])
        singles.append(t)
    for f in range(1, 4):
        states = ms.step(frames[f])
        for s in range(S):
            ref = singles[s].step(frames[f, s])
            assert torch.equal(ms.particles[s], singles[s].particles)      # bit-identical filter
            assert torch.equal(ms.idx[s], singles[s].idx)
            assert np.allclose(states[s].bbox, ref.bbox, atol=1e-9) and states[s].frame_idx == f


def test_fp16_operand_mode_is_closer_to_the_fp32_oracle():
    """dtype=torch.float16 (same tensor-core rate and bytes as bf16, three more mantissa bits):
    the encoder lands ~8x closer to the fp32 oracle, which brings the normalised weights near the
    north-star's 1e-3 figure; bf16 stays the default."""
    from vitparticlefiltertracker_b200 import ops, pack_for_device, Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    from oracle import OracleTracker
    dev = cuda()
    model, n = "vit_tiny_patch16_224", 6
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    rng = np.random.default_rng(11)
    frame = rng.integers(0, 256, size=(200, 260, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(40, 220, n), rng.uniform(40, 160, n),
                      rng.uniform(30, 90, n), rng.uniform(30, 90, n)], 1).astype(np.float32)
    crops = crop_resize(frame, boxes)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    rels = {}
    for dt in (torch.bfloat16, torch.float16):
        enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, sd, dt),
                             dev, dtype=dt)
        patches = ops.crop_boxes(torch.from_numpy(frame).to(dev), torch.from_numpy(boxes).to(dev), dtype=dt)
        ref = torch.from_numpy(patchify(crops)).to(dt).float()
        assert (patches.float().cpu() - ref).abs().max() <= 2.0 ** -7
        got = enc.forward(patches, n).cpu()
        rels[dt] = float((got - want).norm() / want.norm())
    print(f"embedding rel-L2: bf16 {rels[torch.bfloat16]:.3e}, fp16 {rels[torch.float16]:.3e}")
    assert rels[torch.float16] < 2e-3 and rels[torch.float16] < 0.3 * rels[torch.bfloat16]
    # whole tracker, teacher-forced one frame: weights against the oracle
    frames, gt = make_clip("box", 224, 224, 3, seed=0, target=(48, 48))
    ora = OracleTracker(model, 128, seed=0)
    trk = Tracker(model, 128, seed=0, dtype=torch.float16)
    ora.init(frames[0], gt[0]); trk.init(frames[0], gt[0])
    trk.particles.copy_(torch.from_numpy(ora.particles))
    ora.step(frames[1]); trk.step(frames[1])
    dcos = np.abs(trk.logw.cpu().numpy() * np.float32(0.1) - ora.last["cos"]).max()
    relw = np.abs(trk.weights.cpu().numpy() / ora.last["w"] - 1).max()
    print(f"fp16 tracker: |dcos| {dcos:.2e}, weights rel {relw:.2e}")
    # max over 128 particles of a rounding-noise quantity: 4e-4 .. 8e-4 from build to build
    assert dcos < 1e-3 and relw < 1e-2


def test_multistream_crop64_matches_single_stream_trackers():
    """BASELINE configs[4] shape at the native-resolution crop size: S streams batched through one
    encoder call must follow S independent Trackers bit for bit."""
    from vitparticlefiltertracker_b200 import MultiStreamTracker, Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    dev = cuda()
    S, n, model = 3, 96, "vit_tiny_patch16_224"
    clips = [make_clip("needle", 120, 160, 4, seed=s, target=(3, 60)) for s in range(S)]
    frames = np.stack([c[0] for c in clips], axis=1)                        # [F,S,H,W,3]
    mst = MultiStreamTracker(S, model, n, seed=0, crop_size=64)
    mst.init(frames[0], [c[1][0] for c in clips])
    singles = [Tracker(model, n, seed=0, stream_id=s, crop_size=64) for s in range(S)]
    for s, t in enumerate(singles):
        t.init(frames[0, s], clips[s][1][0])
    for f in range(1, 4):
        got = mst.step(frames[f])
        for s, t in enumerate(singles):
            want = t.step(frames[f, s])
            assert torch.equal(mst.particles[s], t.particles), (f, s)
            assert np.allclose(got[s].bbox, want.bbox, atol=1e-9) and got[s].frame_idx == f


def test_main_py_runs_synthetic_and_image_sequence_inputs(tmp_path):
    """The README's `python main.py` usage (README.md:34-42) over the two kinds of input it names:
    a (synthetic) video feed and images on disk; both must give the same track."""
    import json
    import os
    import subprocess
    import sys
    import yaml
    cv2 = pytest.importorskip("cv2")
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    syn = dict(clip="box", height=120, width=160, frames=5, target=[32, 24], seed=3)
    base = dict(model="vit_tiny_patch16_224", n_particles=64, seed=1, crop_size=64, cuda_graph=False)
    frames, gt = make_clip("box", 120, 160, 5, 3, (32, 24))
    for i, fr in enumerate(frames):
        cv2.imwrite(str(tmp_path / f"f{i:02d}.png"), fr[..., ::-1])
    cfgs = {"syn": dict(base, input=dict(kind="synthetic", synthetic=syn)),
            "img": dict(base, input=dict(kind="images", path=str(tmp_path / "f*.png"),
                                         bbox=[float(v) for v in gt[0]]))}
    out = {}
    for name, cfg in cfgs.items():
        p = tmp_path / f"{name}.yaml"
        p.write_text(yaml.safe_dump(cfg))
        r = subprocess.run([sys.executable, os.path.join(root, "main.py"), str(p)], capture_output=True,
                           text=True, timeout=300, cwd=root)
        assert r.returncode == 0, r.stderr[-1500:]
        out[name] = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert [o["frame"] for o in out[name]] == [1, 2, 3, 4]
    assert [o["bbox"] for o in out["syn"]] == [o["bbox"] for o in out["img"]]


def test_full_size_config5_share_matches_a_single_stream_tracker():
    """BASELINE configs[4] per-GPU share at full size: 8 needle streams x 2048 particles, ViT-B/16,
    one batched encoder call over 16 384 crops; stream 5 must follow an independent 2048-particle
    Tracker bit for bit (batch invariance of every kernel at the largest single-GPU batch used)."""
    from vitparticlefiltertracker_b200 import MultiStreamTracker, Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    S, n, model, probe = 8, 2048, "vit_base_patch16_224", 5
    clips = [make_clip("needle", 480, 640, 3, seed=s, target=(3, 120), velocity=(1.5, 0.5)) for s in range(S)]
    frames = np.stack([c[0] for c in clips], axis=1)                        # [F,S,H,W,3]
    mst = MultiStreamTracker(S, model, n, seed=0)
    mst.init(frames[0], [c[1][0] for c in clips])
    one = Tracker(model, n, seed=0, stream_id=probe, encoder=None)
    one.init(frames[0, probe], clips[probe][1][0])
    for f in (1, 2):
        states = mst.step(frames[f])
        ref = one.step(frames[f, probe])
        assert torch.equal(mst.particles[probe], one.particles), f
        assert torch.equal(mst.idx[probe], one.idx)
        assert np.allclose(states[probe].bbox, ref.bbox, atol=1e-9)
        assert all(np.isfinite(st.bbox).all() and 1.0 <= st.ess <= n for st in states)
