"""GPU parity of the optional fp16 residual stream (VPF_RESIDUAL_F16): the
residual epilogues (in place through the TMA unit for K <= 1536, register-staged above), the
patch-embed row remap, LayerNorm on an fp16 stream, and the whole encoder / tracker against the
fp32 oracle at the same stated tolerances as the default fp32 stream."""
import numpy as np
import pytest
import torch

from oracle import OracleTracker, pf
from oracle.crop import crop_resize, patchify
from oracle.vit import VIT_CONFIGS, init_vit_weights, vit_forward
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("M,N,K", [(197 * 5, 768, 768), (197 * 5, 768, 3072), (300, 192, 192),
                                   (1000, 1024, 4096), (77, 384, 1536)])
def test_residual_epilogues_fp16_stream(M, N, K):
    """K <= 1536 takes the TMA in-place epilogue, K > 1536 the register-staged one."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    x0 = (torch.randn(M, N, generator=g) * 3).to(torch.float16)
    ref = x0.float().to(dev) + a.float() @ w.float().t() + bias
    arena = torch.full((M * N + 8192,), 9.0, dtype=torch.float16, device=dev)
    x = arena[4096:4096 + M * N].view(M, N)
    x.copy_(x0)
    ops.gemm_bf16(a, w, bias, 2, out=x)
    torch.cuda.synchronize()
    assert bool((arena[:4096] == 9.0).all()) and bool((arena[4096 + M * N:] == 9.0).all())
    err = (x.float() - ref).abs()
    assert bool((err <= 2.0 ** -10 * ref.abs() + 2e-3).all()), float(err.max())   # fp16 rounding of the sum
    # saturation instead of inf
    x.fill_(60000.0)
    big_bias = torch.full((N,), 30000.0, device=dev)
    ops.gemm_bf16(a, w, big_bias, 2, out=x)
    assert torch.isfinite(x).all() and float(x.max()) == 65504.0


def test_patch_embed_and_layernorm_fp16_stream():
    from vitparticlefiltertracker_b200 import _lib, ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    n, D, P = 5, 192, 196
    a = torch.randn(n * P, 768, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(D, 768, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    pos = torch.randn(P + 1, D, generator=g).to(dev)
    X = torch.full((n * (P + 1), D), 7.0, dtype=torch.float16, device=dev)
    ops.gemm_bf16(a, w, bias, 3, out=X, pos=pos)
    ref = (a.float() @ w.float().t() + bias).view(n, P, D) + pos[1:][None]
    Xv = X.float().view(n, P + 1, D)
    assert (Xv[:, 1:] - ref).abs().max() < 6e-3
    assert (Xv[:, 0] == 7.0).all()


@pytest.mark.parametrize("model,n,S", [("vit_tiny_patch16_224", 9, 224), ("vit_base_patch16_224", 3, 224),
                                       ("vit_tiny_patch16_224", 6, 64), ("vit_large_patch16_224", 2, 224)])
def test_encoder_fp16_residual_matches_oracle(model, n, S):
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    g = torch.Generator().manual_seed(7)
    for k in sd:                                  # non-trivial biases: exercise every fused term
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    rng = np.random.default_rng(n)
    frame = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(40, 280, n), rng.uniform(40, 200, n),
                      rng.uniform(30, 90, n), rng.uniform(30, 90, n)], 1).astype(np.float32)
    crops = crop_resize(frame, boxes, size=S)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    packed = pack_for_device(model, sd, crop_size=S)
    patches = torch.from_numpy(patchify(crops)).to(torch.bfloat16).to(dev)
    out = {}
    for rd in (torch.float32, torch.float16):
        enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, packed, dev, crop_size=S,
                             residual_dtype=rd)
        out[rd] = enc.forward(patches, n).cpu()
        rel = _rel(out[rd], want)
        cos = torch.nn.functional.cosine_similarity(out[rd], want, dim=1).min()
        print(f"{model} S={S} residual {rd}: rel-L2 {rel:.3e}, min cos {cos:.6f}")
        assert rel < 1.5e-2 and cos > 1 - 1e-3
    # the fp16 stream adds at most a small fraction to the distance from the oracle
    assert _rel(out[torch.float16], want) < _rel(out[torch.float32], want) * 1.25 + 5e-4


def test_tracker_fp16_residual_teacher_forced_and_graph():
    from vitparticlefiltertracker_b200 import Tracker
    from vitparticlefiltertracker_b200.synthetic import make_clip
    cuda()
    frames, gt = make_clip("box", 224, 224, 8, seed=0, target=(48, 48))
    n = 256
    ora = OracleTracker("vit_tiny_patch16_224", n, seed=0)
    trk = Tracker("vit_tiny_patch16_224", n, seed=0, residual_dtype=torch.float16)
    ora.init(frames[0], gt[0])
    trk.init(frames[0], gt[0])
    for t in range(1, 4):
        trk.particles.copy_(torch.from_numpy(ora.particles))
        so, sg = ora.step(frames[t]), trk.step(frames[t])
        dcos = np.abs(trk.logw.cpu().numpy() * np.float32(0.1) - ora.last["cos"]).max()
        print(f"fp16 residual frame {t}: |dcos| {dcos:.2e}  ess {sg.ess:.1f}/{so.ess:.1f}")
        assert dcos < 5e-3 and np.abs(np.array(sg.bbox) - np.array(so.bbox)).max() < 0.5
        assert np.array_equal(trk.idx.cpu().numpy(),
                              pf.systematic_resample(trk.weights.cpu().numpy(), trk.last_u))
    eager = Tracker("vit_tiny_patch16_224", 128, seed=5, residual_dtype=torch.float16)
    graph = Tracker("vit_tiny_patch16_224", 128, seed=5, residual_dtype=torch.float16, cuda_graph=True)
    eager.init(frames[0], gt[0]); graph.init(frames[0], gt[0])
    for t in range(1, 5):
        assert eager.step(frames[t]) == graph.step(frames[t])
        assert torch.equal(eager.particles, graph.particles)
    with pytest.raises(ValueError):
        Tracker("vit_tiny_patch16_224", 16, residual_dtype=torch.bfloat16)
