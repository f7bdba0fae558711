"""GPU parity: tcgen05 GEMM epilogues, LayerNorm, attention and the full encoder."""
import numpy as np
import pytest
import torch

from oracle.vit import init_vit_weights, vit_forward, VIT_CONFIGS
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 768), (1970, 2304, 768),
                                   (777, 576, 192), (513, 384, 1536), (1024, 128, 128),
                                   (197 * 64, 3072, 768)])
def test_gemm_bias_bf16(M, N, K):
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm_bf16(a, w, bias, 0).float()
    assert _rel(out, ref) < 4e-3, _rel(out, ref)               # bf16 output rounding only
    assert (out - ref).abs().max() <= 0.02 * ref.abs().max()
    # run twice back to back: TMEM / mbarrier state must be clean between launches
    out2 = ops.gemm_bf16(a, w, bias, 0).float()
    assert torch.equal(out, out2)


def test_gemm_gelu():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(1)
    M, N, K = 1000, 3072, 768
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.06).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    out = ops.gemm_bf16(a, w, bias, 1).float()
    assert ((out - ref).abs() <= 0.01 * ref.abs() + 2e-3).all() and _rel(out, ref) < 4e-3


def test_gelu_epilogue_accuracy_against_fp64_erf_gelu():
    """The epilogue's GELU alone (identity weights, zero bias): every output within the bf16
    rounding of the exact erf-GELU plus the stated error of the tanh form (2.5e-5 formula error,
    2^-11 relative on tanh.approx => <= 2.4e-4 |x|); rms error far below the rounding's."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(4)
    M, N = 4096, 128
    x = torch.cat([torch.randn(M - 64, N, generator=g) * 1.5,
                   torch.linspace(-9, 9, 64 * N).view(64, N)]).to(torch.bfloat16)
    out = ops.gemm_bf16(x.to(dev), torch.eye(N).to(torch.bfloat16).to(dev), torch.zeros(N, device=dev), 1)
    xd = x.double()
    ref = 0.5 * xd * (1 + torch.erf(xd / 2 ** 0.5))
    err = (out.cpu().double() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 3e-4 * xd.abs() + 4e-5).all()), float(err.max())
    exact_rounded = ref.to(torch.bfloat16).double()
    extra = (out.cpu().double() - exact_rounded).abs()            # beyond the unavoidable rounding
    rms_extra = float(extra.pow(2).mean().sqrt())
    rms_round = float((exact_rounded - ref).pow(2).mean().sqrt())
    print(f"gelu epilogue: rms beyond rounding {rms_extra:.2e}, rounding itself {rms_round:.2e}")
    assert rms_extra < 0.6 * rms_round + 1e-5


def test_gemm_residual_and_patch_embed():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(2)
    M, N, K = 197 * 5, 768, 3072
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    ref = x + a.float() @ w.float().t() + bias
    ops.gemm_bf16(a, w, bias, 2, out=x)
    assert (x - ref).abs().max() < 2e-3, (x - ref).abs().max()  # fp32 out: accumulation order only
    # patch embed: rows remapped into the 197-token stream, + pos
    n, D = 3, 192
    a = torch.randn(n * 196, 768, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(D, 768, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    pos = torch.randn(197, D, generator=g).to(dev)
    X = torch.full((n * 197, D), 7.0, device=dev)
    ops.gemm_bf16(a, w, bias, 3, out=X, pos=pos)
    ref = (a.float() @ w.float().t() + bias).view(n, 196, D) + pos[1:][None]
    Xv = X.view(n, 197, D)
    assert (Xv[:, 1:] - ref).abs().max() < 2e-3
    assert (Xv[:, 0] == 7.0).all()                              # CLS rows untouched by the GEMM


def test_layernorm():
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    for D in (192, 768, 1024):
        x = (torch.randn(999, D, generator=g) * 3 + 1).to(dev)
        gm = (1 + 0.1 * torch.randn(D, generator=g)).to(dev)
        bt = (0.1 * torch.randn(D, generator=g)).to(dev)
        ref = torch.nn.functional.layer_norm(x, (D,), gm, bt, 1e-6)
        out = ops.layernorm(x, gm, bt).float()
        assert (out - ref).abs().max() < 0.04 and _rel(out, ref) < 4e-3


@pytest.mark.parametrize("heads,n", [(3, 5), (12, 3)])
def test_attention(heads, n):
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    D = heads * 64
    g = torch.Generator(device="cpu").manual_seed(heads)
    qkv = (torch.randn(n * 197, 3 * D, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    out = ops.attention(qkv, n, heads).float()
    q, k, v = qkv.float().view(n, 197, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(n * 197, D)
    assert (out - ref).abs().max() < 0.03, (out - ref).abs().max()
    assert _rel(out, ref) < 8e-3, _rel(out, ref)


@pytest.mark.parametrize("model,n", [("vit_tiny_patch16_224", 9), ("vit_base_patch16_224", 3)])
def test_vit_forward_matches_oracle(model, n):
    """End-to-end embedding vs the fp32 CPU oracle (bf16 operands, fp32 accumulate/residual).
    Stated bar (D12): rel-L2 <= 1e-2, cosine between the two embeddings >= 1 - 1e-3."""
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    from oracle.crop import crop_resize, patchify
    dev = cuda()
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(40, 280, n), rng.uniform(40, 200, n),
                      rng.uniform(30, 90, n), rng.uniform(30, 90, n)], 1).astype(np.float32)
    crops = crop_resize(frame, boxes)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, sd), dev)
    patches = torch.from_numpy(patchify(crops)).to(torch.bfloat16).to(dev)
    got = enc.forward(patches, n).cpu()
    rel = _rel(got, want)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    print(f"{model}: rel-L2 {rel:.3e}, min cos {cos.min():.6f}")
    assert rel < 1e-2, rel
    assert cos.min() > 1 - 1e-3
    # a smaller batch through the same handle gives bit-identical rows (batch invariance)
    got2 = enc.forward(patches[:196 * 2].contiguous(), 2).cpu()
    assert torch.equal(got2, got[:2])


def test_cls_only_last_block_matches_full_last_block(monkeypatch):
    """The CLS-only last block is exact pruning: same embedding as running the last block on
    every token (up to the bf16 rounding of the one attention row computed by another kernel)."""
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    model, n = "vit_tiny_patch16_224", 5
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    g = torch.Generator(device="cpu").manual_seed(5)
    patches = torch.randn(n * 196, 768, generator=g).to(torch.bfloat16).to(dev)
    packed = pack_for_device(model, sd)
    pruned = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, packed, dev)
    a = pruned.forward(patches, n).cpu()
    monkeypatch.setenv("VPF_CLS_PRUNE", "0")
    full = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, packed, dev)
    assert full.launches == 2 + 7 * 12 + 1 and pruned.launches == 2 + 7 * 11 + 10
    b = full.forward(patches, n).cpu()
    assert _rel(a, b) < 2e-3, _rel(a, b)


def test_vit_forward_with_nontrivial_biases_and_layernorm_params():
    """The seeded init has zero biases and unit LayerNorm gains; a real (timm) checkpoint does
    not.  Randomise every bias / gamma / beta / cls token so that each fused term is exercised
    end to end, through the same state_dict= path a pretrained checkpoint would take."""
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    dev = cuda()
    model, n = "vit_tiny_patch16_224", 7
    cfg = VIT_CONFIGS[model]
    g = torch.Generator().manual_seed(3)
    sd = {}
    for k, v in init_vit_weights(model).items():
        if k.endswith("bias"):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1 + torch.randn(v.shape, generator=g) * 0.2
        elif k == "cls_token":
            sd[k] = torch.randn(v.shape, generator=g) * 0.05
        else:
            sd[k] = v * 2.0
    crops = torch.randn(n, 3, 224, 224, generator=g).clamp(-1, 1)
    want = torch.from_numpy(vit_forward(crops.numpy(), sd, model))
    from oracle.crop import patchify
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, sd), dev)
    got = enc.forward(torch.from_numpy(patchify(crops.numpy())).to(torch.bfloat16).to(dev), n).cpu()
    rel = _rel(got, want)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    print(f"randomised checkpoint: rel-L2 {rel:.3e}, min cos {cos.min():.6f}")
    assert rel < 1.5e-2 and cos.min() > 1 - 1e-3


def test_in_step_kernel_timing_hook():
    """vpf_vit_set_timing / vpf_vit_timing (bench.py's encoder_breakdown_ms): an event pair around
    every launch of a forward; the per-kind sums cover the forward and change nothing in its result."""
    from vitparticlefiltertracker_b200 import ops, init_state_dict, pack_for_device, VIT_CONFIGS
    dev = cuda()
    model, n = "vit_tiny_patch16_224", 64
    cfg = VIT_CONFIGS[model]
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, init_state_dict(model)), dev)
    patches = (torch.randn(n * 196, 768, device=dev) * 0.5).to(torch.bfloat16)
    ref = enc.forward(patches, n).clone()
    assert all(v == 0.0 for v in enc.timing().values())             # off: nothing recorded
    enc.set_timing(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    got = enc.forward(patches, n)
    e.record()
    t = enc.timing()
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    assert set(t) == set(ops.VitEncoder.TIMING_KINDS) and all(v > 0 for v in t.values())
    assert 0.5 * s.elapsed_time(e) <= sum(t.values()) <= 1.05 * s.elapsed_time(e)
    enc.set_timing(False)
    assert torch.equal(enc.forward(patches, n), ref)
