"""GPU parity of the variable crop size path ("native-resolution crops":
S x S crops, (S/16)^2 + 1 tokens) against the oracle and the committed golden vectors
tests/golden/oracle_v2_crop.npz.  Covers the crop kernel at every grid parity, the three attention
kernels that serve T < 197 (one warp per pair for T <= 32 and T <= 64, CTA per pair above), the
patch-embed row remap, the encoder, and the whole Tracker."""
import os

import numpy as np
import pytest
import torch

from oracle import OracleTracker, pf
from oracle.crop import boxes_from_state, crop_resize, patchify
from oracle.vit import VIT_CONFIGS, init_vit_weights, vit_forward
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def _bf16(x):
    return torch.from_numpy(x).to(torch.bfloat16).float().numpy()


def _rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "oracle_v2_crop.npz"))


@pytest.mark.parametrize("S", [16, 48, 64, 112, 176, 208])
def test_crop_sizes_match_oracle(S):
    """Odd and even patch-row counts (the last CTA of a crop handles one patch row when S/16 is odd)."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    H, W = 360, 500
    rng = np.random.default_rng(S)
    frame = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    boxes = np.array([[W / 2, H / 2, 64, 64], [10, 10, 50, 80], [W - 5, H - 5, 96, 96],
                      [W / 3 + 0.37, H / 3 + 0.81, 17.3, 9.1], [-300, -300, 20, 20],
                      [W / 2, H / 2, W, H], [W / 2, H / 2, S, S]], dtype=np.float32)
    want = _bf16(patchify(crop_resize(frame, boxes, size=S)))
    got = ops.crop_boxes(torch.from_numpy(frame).to(dev), torch.from_numpy(boxes).to(dev), crop_size=S)
    got = got.float().cpu().numpy()
    assert got.shape == (len(boxes) * (S // 16) ** 2, 768)
    diff = np.abs(got - want)
    assert diff.max() <= 2.0 ** -7
    assert (diff > 0).mean() < 1e-3
    # state form
    n = 21
    st = np.zeros((6, n), dtype=np.float32)
    st[0] = rng.uniform(0, W, n); st[1] = rng.uniform(0, H, n); st[4] = rng.normal(0, 0.2, n)
    want = _bf16(patchify(crop_resize(frame, boxes_from_state(st, 48, 32), size=S)))
    got = ops.crop_state(torch.from_numpy(frame).to(dev), torch.from_numpy(st).to(dev), 48.0, 32.0,
                         crop_size=S)
    assert np.abs(got.float().cpu().numpy() - want).max() <= 2.0 ** -7


def test_bad_crop_size_is_rejected():
    from vitparticlefiltertracker_b200 import ops, Tracker
    dev = cuda()
    frame = torch.zeros((64, 64, 3), dtype=torch.uint8, device=dev)
    boxes = torch.tensor([[32.0, 32.0, 16.0, 16.0]], device=dev)
    with pytest.raises(ValueError):
        ops.crop_boxes(frame, boxes, crop_size=100)
    with pytest.raises(ValueError):
        Tracker("vit_tiny_patch16_224", 16, crop_size=240)


@pytest.mark.parametrize("tokens", [2, 5, 10, 17, 26, 32, 33, 37, 50, 64, 65, 101, 112, 113, 145, 170, 196])
@pytest.mark.parametrize("heads,n", [(3, 7)])
def test_attention_token_counts(tokens, heads, n):
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    D = heads * 64
    g = torch.Generator(device="cpu").manual_seed(tokens)
    qkv = (torch.randn(n * tokens, 3 * D, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    out = ops.attention(qkv, n, heads, tokens).float()
    q, k, v = qkv.float().view(n, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(n * tokens, D)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 0.03, (out - ref).abs().max()
    assert _rel(out, ref) < 8e-3, _rel(out, ref)


def test_attention_small_many_pairs():
    """More (crop, head) pairs than one wave, pair count not a multiple of the warps per CTA."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    heads, n, tokens = 12, 601, 17
    D = heads * 64
    g = torch.Generator(device="cpu").manual_seed(0)
    qkv = torch.randn(n * tokens, 3 * D, generator=g).to(torch.bfloat16).to(dev)
    out = ops.attention(qkv, n, heads, tokens).float()
    q, k, v = qkv.float().view(n, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).transpose(1, 2).reshape(n * tokens, D)
    assert _rel(out, ref) < 8e-3


def test_patch_embed_row_remap_small_grid():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(2)
    n, D, P = 37, 192, 16
    a = torch.randn(n * P, 768, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(D, 768, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    pos = torch.randn(P + 1, D, generator=g).to(dev)
    X = torch.full((n * (P + 1), D), 7.0, device=dev)
    ops.gemm_bf16(a, w, bias, 3, out=X, pos=pos, patches_per_crop=P)
    ref = (a.float() @ w.float().t() + bias).view(n, P, D) + pos[1:][None]
    Xv = X.view(n, P + 1, D)
    assert (Xv[:, 1:] - ref).abs().max() < 2e-3
    assert (Xv[:, 0] == 7.0).all()


@pytest.mark.parametrize("model,n,S", [("vit_tiny_patch16_224", 9, 64), ("vit_tiny_patch16_224", 1, 16),
                                       ("vit_tiny_patch16_224", 5, 112), ("vit_tiny_patch16_224", 4, 128),
                                       ("vit_base_patch16_224", 3, 64), ("vit_large_patch16_224", 2, 96),
                                       ("vit_small_patch16_224", 4, 224), ("vit_small_patch16_224", 3, 80)])
def test_vit_forward_small_crops_matches_oracle(model, n, S):
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    rng = np.random.default_rng(S + n)
    frame = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(40, 280, n), rng.uniform(40, 200, n),
                      rng.uniform(30, 90, n), rng.uniform(30, 90, n)], 1).astype(np.float32)
    crops = crop_resize(frame, boxes, size=S)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n,
                         pack_for_device(model, sd, crop_size=S), dev, crop_size=S)
    patches = torch.from_numpy(patchify(crops)).to(torch.bfloat16).to(dev)
    got = enc.forward(patches, n).cpu()
    rel = _rel(got, want)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    print(f"{model} S={S}: rel-L2 {rel:.3e}, min cos {cos.min():.6f}")
    assert rel < 1.5e-2, rel
    assert cos.min() > 1 - 1e-3
    if n > 1:   # batch invariance through the same handle
        P = (S // 16) ** 2
        got1 = enc.forward(patches[:P].contiguous(), 1).cpu()
        assert torch.equal(got1, got[:1])


def test_encoder_rejects_pos_embed_of_another_crop_size():
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    model = "vit_tiny_patch16_224"
    cfg = VIT_CONFIGS[model]
    with pytest.raises(ValueError):
        ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], 2,
                       pack_for_device(model, init_vit_weights(model)), dev, crop_size=64)


def test_crop64_golden(G):
    from vitparticlefiltertracker_b200 import ops, init_state_dict, pack_for_device
    dev = cuda()
    model = "vit_tiny_patch16_224"
    cfg = VIT_CONFIGS[model]
    frame = torch.from_numpy(G["crop_frame"]).to(dev)
    boxes = torch.from_numpy(G["crop_boxes"]).to(dev)
    sd = init_state_dict(model)
    for S in (64, 112):
        patches = ops.crop_boxes(frame, boxes, crop_size=S)
        p = patches.float().cpu().numpy()
        assert np.abs(p[::3, ::41] - _bf16(G[f"patch_samples_{S}"])).max() <= 2.0 ** -7
        packed = pack_for_device(model, sd, crop_size=S)
        assert np.allclose(packed[3][:, :8].numpy(), G[f"pos_embed_{S}"], atol=1e-6)
        enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], 4, packed, dev, crop_size=S)
        emb = enc.forward(patches, 4).cpu()
        want = torch.from_numpy(G[f"vit_emb_{S}"])
        cos = torch.nn.functional.cosine_similarity(emb, want, dim=1).min()
        assert _rel(emb, want) < 1e-2 and cos > 1 - 1e-3, (S, _rel(emb, want), float(cos))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_tracker_crop64_teacher_forced(dtype):
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    frames, gt = make_clip("box", 224, 224, 12, seed=0, target=(48, 48))
    n = 256
    ora = OracleTracker("vit_tiny_patch16_224", n, seed=0, crop_size=64)
    trk = Tracker("vit_tiny_patch16_224", n, seed=0, crop_size=64, dtype=dtype)
    ora.init(frames[0], gt[0])
    trk.init(frames[0], gt[0])
    tm = torch.nn.functional.cosine_similarity(trk.template.cpu(), torch.from_numpy(ora.template), 0)
    assert tm > 1 - 1e-3
    for t in range(1, 4):
        trk.particles.copy_(torch.from_numpy(ora.particles))
        so = ora.step(frames[t])
        sg = trk.step(frames[t])
        assert sg.frame_idx == so.frame_idx == t and trk.last_u == ora.last["u"]
        assert np.abs(trk.propagated.cpu().numpy() - ora.last["propagated"]).max() < 2e-5
        dcos = np.abs(trk.logw.cpu().numpy() * np.float32(0.1) - ora.last["cos"]).max()
        relw = np.abs(trk.weights.cpu().numpy() / ora.last["w"] - 1).max()
        print(f"S=64 {dtype} frame {t}: |dcos| {dcos:.2e}  weights rel {relw:.2e}  ess {sg.ess:.1f}/{so.ess:.1f}")
        assert dcos < 5e-3
        assert np.abs(np.array(sg.bbox) - np.array(so.bbox)).max() < 0.5
        assert abs(sg.ess - so.ess) / so.ess < 0.15
        assert np.array_equal(trk.idx.cpu().numpy(),
                              pf.systematic_resample(trk.weights.cpu().numpy(), trk.last_u))
    for t in range(4, 12):
        s = trk.step(frames[t])
        assert np.isfinite(s.bbox).all() and 1 <= s.ess <= n


def test_tracker_crop64_cuda_graph_and_checkpoint():
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    frames, gt = make_clip("box", 224, 224, 8, seed=1, target=(48, 48))
    eager = Tracker("vit_tiny_patch16_224", 200, seed=5, crop_size=64)
    graph = Tracker("vit_tiny_patch16_224", 200, seed=5, crop_size=64, cuda_graph=True)
    eager.init(frames[0], gt[0])
    graph.init(frames[0], gt[0])
    for t in range(1, 5):
        a, b = eager.step(frames[t]), graph.step(frames[t])
        assert torch.equal(eager.particles, graph.particles) and a == b
    ckpt = eager.state_dict()
    assert ckpt["crop_size"] == 64
    other = Tracker("vit_tiny_patch16_224", 200, seed=5)            # default crop size: must refuse
    with pytest.raises(ValueError):
        other.load_state_dict(ckpt)


def test_full_size_vit_base_4096_crop64_properties():
    """The 'native-resolution' operating point of BASELINE configs[2]: 4096 particles, ViT-B/16,
    64-pixel crops (17 tokens).  Size-independent properties + batch invariance + sampled parity."""
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    dev = cuda()
    frames, gt = make_clip("box", 1080, 1920, 3, seed=0, target=(96, 96))
    n, S = 4096, 64
    trk = Tracker("vit_base_patch16_224", n, seed=0, crop_size=S)
    trk.init(frames[0], gt[0])
    s1 = trk.step(frames[1])
    assert s1.frame_idx == 1 and 1.0 <= s1.ess <= n and np.isfinite(s1.bbox).all()
    w = trk.weights.cpu().numpy()
    assert abs(float(w.astype(np.float64).sum()) - 1.0) < 1e-4 and (w >= 0).all()
    idx = trk.idx.cpu().numpy()
    assert np.array_equal(idx, pf.systematic_resample(w, trk.last_u))
    assert torch.equal(trk.particles, trk.propagated[:, trk.idx.long()])
    st = trk.propagated.cpu().numpy()
    sel = np.arange(0, n, 512)
    crops = crop_resize(frames[1], boxes_from_state(st[:, sel], trk.w0, trk.h0), size=S)
    got = trk.patches.view(n, 16, 768)[torch.from_numpy(sel).to(dev)].reshape(-1, 768).float().cpu().numpy()
    assert np.abs(got - _bf16(patchify(crops))).max() <= 2.0 ** -7
    # embeddings of the sampled particles vs the fp32 oracle, and batch invariance (N=4096 vs N=8)
    want = torch.from_numpy(vit_forward(crops, init_vit_weights("vit_base_patch16_224"), "vit_base_patch16_224"))
    emb = trk.emb[torch.from_numpy(sel).to(dev)].cpu()
    cos = torch.nn.functional.cosine_similarity(emb, want, dim=1).min()
    assert _rel(emb, want) < 1e-2 and cos > 1 - 1e-3
    small = trk.encoder.forward(trk.patches.view(n, 16 * 768)[torch.from_numpy(sel).to(dev)].reshape(-1, 768).contiguous(), len(sel)).cpu()
    assert torch.equal(small, emb)


def test_no_out_of_bounds_writes_with_canaries():
    """compute-sanitizer is closed on this pool, so writes are fenced by hand: every output lives in
    the middle of a sentinel-filled arena, and the sentinel on both sides must survive."""
    from vitparticlefiltertracker_b200 import _lib, ops
    dev = cuda()
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    PAD = 4096
    rng = np.random.default_rng(0)
    frame = torch.from_numpy(rng.integers(0, 256, size=(97, 131, 3), dtype=np.uint8)).to(dev)

    def arena(n_elems, dtype, fill):
        a = torch.full((n_elems + 2 * PAD,), fill, dtype=dtype, device=dev)
        return a, a[PAD:PAD + n_elems]

    def intact(a, n_elems, fill):
        return bool((a[:PAD] == fill).all()) and bool((a[PAD + n_elems:] == fill).all())

    # crop, every grid parity, boxes that up-sample, down-sample and hang over the frame
    boxes = torch.tensor([[60, 40, 30, 30], [5, 5, 200, 150], [128, 95, 64, 64], [60, 40, 9, 300]],
                         dtype=torch.float32, device=dev)
    for S in (16, 48, 64, 112, 224):
        n_el = boxes.shape[0] * (S // 16) ** 2 * 768
        a, out = arena(n_el, torch.bfloat16, 7.0)
        _lib.check(lib.vpf_crop_boxes(frame.data_ptr(), 97, 131, boxes.data_ptr(), boxes.shape[0],
                                      out.data_ptr(), _lib.DTYPE_BF16, S, stream), "vpf_crop_boxes")
        torch.cuda.synchronize()
        assert intact(a, n_el, 7.0), S
        assert bool((out.float().abs() <= 1.0).all()), S          # every element was written
    # attention: all three small-token kernels
    heads, n = 3, 5
    for tokens in (2, 17, 33, 64, 65, 150):
        qkv = torch.randn(n * tokens, 3 * 64 * heads, device=dev).to(torch.bfloat16)
        n_el = n * tokens * 64 * heads
        a, out = arena(n_el, torch.bfloat16, 1000.0)
        _lib.check(lib.vpf_attention(qkv.data_ptr(), out.data_ptr(), n, heads, tokens, stream),
                   "vpf_attention")
        torch.cuda.synchronize()
        assert intact(a, n_el, 1000.0), tokens
        assert bool((out.float().abs() < 100.0).all()), tokens
    # patch-embed row remap (fp32 residual stream), 16 patches per crop
    ncrop, D, P = 9, 192, 16
    A = torch.randn(ncrop * P, 768, device=dev).to(torch.bfloat16)
    Wt = (torch.randn(D, 768, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(D, device=dev)
    pos = torch.zeros(P + 1, D, device=dev)
    n_el = ncrop * (P + 1) * D
    a, X = arena(n_el, torch.float32, 55.0)
    _lib.check(lib.vpf_gemm_bf16(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), X.data_ptr(),
                                 pos.data_ptr(), ncrop * P, D, 768, 3, P, 0, stream), "vpf_gemm_bf16")
    torch.cuda.synchronize()
    assert intact(a, n_el, 55.0)
    assert bool((X.view(ncrop, P + 1, D)[:, 0] == 55.0).all())     # CLS rows are not the GEMM's
    # whole encoder at 17 tokens: workspace and embedding output fenced
    from vitparticlefiltertracker_b200 import pack_for_device, init_state_dict
    model = "vit_tiny_patch16_224"
    cfg = VIT_CONFIGS[model]
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], 6,
                         pack_for_device(model, init_state_dict(model), crop_size=64), dev, crop_size=64)
    patches = torch.randn(6 * 16, 768, device=dev).to(torch.bfloat16)
    a, emb = arena(6 * cfg["dim"], torch.float32, -3.0)
    ws_tail = enc.workspace[-8:].clone()
    enc.forward(patches, 6, emb.view(6, cfg["dim"]))
    torch.cuda.synchronize()
    assert intact(a, 6 * cfg["dim"], -3.0) and torch.isfinite(emb).all()
    assert torch.equal(enc.workspace[-8:], ws_tail)
