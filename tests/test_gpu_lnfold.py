"""GPU parity of the LayerNorm-folded GEMMs (include/vpf.h `vpf_vit_set_ln_fold`): row statistics,
the folded qkv / fc1 epilogue, the partial sums the residual epilogues leave, and the encoder
with and without folding against the fp32 CPU oracle (oracle/vit.py norm1 -> qkv, norm2 -> fc1)."""
import os

import numpy as np
import pytest
import torch

from oracle.crop import crop_resize, patchify
from oracle.vit import VIT_CONFIGS, init_vit_weights, vit_forward
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _stream_like(M, D, g, dev):
    """An fp16 'residual stream': per-row offset and scale, a few large channels."""
    x = torch.randn(M, D, generator=g) * (0.5 + 2 * torch.rand(M, 1, generator=g))
    x += torch.randn(M, 1, generator=g) * 0.7
    x[:, ::97] *= 8.0
    return x.to(torch.float16).to(dev)


@pytest.mark.parametrize("D", [192, 384, 768, 1024])
def test_rowstats16(D):
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator().manual_seed(D)
    x = _stream_like(1999, D, g, dev)
    st = ops.rowstats16(x)
    xd = x.double()
    mu = xd.mean(1)
    rstd = 1.0 / torch.sqrt(xd.var(1, unbiased=False) + 1e-6)
    assert (st[:, 0].double() - mu).abs().max() < 2e-6 * max(1.0, float(mu.abs().max()))
    assert ((st[:, 1].double() - rstd).abs() / rstd).max() < 5e-6


def _fold(w, bias, gamma, beta):
    wf = w.double() * gamma.double()[None]
    wf = (wf - wf.mean(1, keepdim=True)).float().to(torch.float16)
    return wf, (bias.double() + w.double() @ beta.double()).float()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,gelu", [(300, 768, 768, False), (1970, 2304, 768, False),
                                        (777, 576, 192, False), (1000, 3072, 768, True),
                                        (513, 1536, 384, True), (260, 4096, 1024, True)])
def test_gemm_ln_matches_layernorm_then_linear(M, N, K, gelu, dtype):
    """fp16 stream x fp16 folded weights, output in the encoder's 16-bit type."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator().manual_seed(M + N + K)
    x = _stream_like(M, K, g, dev)
    w = torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g) * 0.1
    gamma = 1 + 0.2 * torch.randn(K, generator=g)
    beta = 0.1 * torch.randn(K, generator=g)
    wf, bf = (t.to(dev) for t in _fold(w, bias, gamma, beta))
    stats = ops.rowstats16(x)
    out = ops.gemm_ln(x, wf, bf, stats, gelu=gelu, dtype=dtype)
    assert out.dtype == dtype
    out = out.float()
    ref = torch.nn.functional.layer_norm(x.float(), (K,), gamma.to(dev), beta.to(dev), 1e-6) \
        @ w.to(dev).t() + bias.to(dev)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    # distance = fp16 rounding of W' and the rounding of the 16-bit output (the normalised
    # activation itself is never rounded)
    tol = 3e-3 if dtype == torch.bfloat16 else 6e-4
    assert _rel(out, ref) < tol, _rel(out, ref)
    assert (out - ref).abs().max() <= 0.02 * ref.abs().max() + 2e-3
    out2 = ops.gemm_ln(x, wf, bf, stats, gelu=gelu, dtype=dtype).float()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("M,N,K", [(197 * 5, 768, 768), (197 * 5, 768, 3072), (333, 192, 192),
                                   (333, 192, 768), (260, 1024, 4096), (260, 384, 384)])
def test_residual_epilogue_partial_sums(M, N, K):
    """K <= 1536 takes the TMA read-modify-write epilogue, larger K the register-staged one: both
    must update the fp16 stream exactly as without statistics and leave (sum x, sum x^2) of the
    stored values per column group."""
    from vitparticlefiltertracker_b200 import ops
    dev = cuda()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.03).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    x0 = _stream_like(M, N, g, dev)
    x_plain = x0.clone()
    ops.gemm_bf16(a, w, bias, 2, out=x_plain)
    x = x0.clone()
    parts = ops.gemm_residual_stats(a, w, bias, x)
    assert torch.equal(x, x_plain)
    P = ops.ln_stats_parts(N)
    assert parts.shape == (P, M, 2)
    xg = x.double().view(M, P, N // P)
    assert (parts[:, :, 0].double().t() - xg.sum(2)).abs().max() < 1e-4 * float(xg.abs().sum(2).max())
    assert ((parts[:, :, 1].double().t() - (xg * xg).sum(2)).abs()
            / (xg * xg).sum(2)).max() < 1e-5
    # the folded GEMM gives the same answer from the partial sums as from the statistics kernel
    wq = torch.randn(256, N, generator=g) * 0.05
    wq = (wq - wq.mean(1, keepdim=True)).to(torch.float16).to(dev)
    bq = torch.randn(256, generator=g).to(dev)
    o_final = ops.gemm_ln(x, wq, bq, ops.rowstats16(x)).float()
    o_parts = ops.gemm_ln(x, wq, bq, parts).float()
    assert _rel(o_parts, o_final) < 1e-4            # at most an output ulp here and there


def _encoders(model, n, S, sd, dev, modes, dtype=torch.bfloat16):
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    from vitparticlefiltertracker_b200.vit_weights import pack_ln_fold
    cfg = VIT_CONFIGS[model]
    packed = pack_for_device(model, sd, dtype, crop_size=S)
    encs = {}
    for name, env in modes.items():
        old = os.environ.get("VPF_LN_FOLD")
        if env is not None:
            os.environ["VPF_LN_FOLD"] = env
        try:
            encs[name] = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, packed, dev, dtype=dtype,
                                        crop_size=S, residual_dtype=torch.float16,
                                        ln_fold=None if name == "plain" else pack_ln_fold(model, sd, dtype))
        finally:
            os.environ.pop("VPF_LN_FOLD", None)
            if old is not None:
                os.environ["VPF_LN_FOLD"] = old
    return encs


@pytest.mark.parametrize("model,n,S,dtype", [
    ("vit_tiny_patch16_224", 9, 224, torch.bfloat16), ("vit_base_patch16_224", 3, 224, torch.bfloat16),
    ("vit_small_patch16_224", 4, 224, torch.bfloat16), ("vit_tiny_patch16_224", 6, 64, torch.bfloat16),
    ("vit_large_patch16_224", 2, 224, torch.bfloat16),
    ("vit_tiny_patch16_224", 24, 224, torch.float16), ("vit_base_patch16_224", 3, 224, torch.float16)])
def test_encoder_ln_fold_matches_oracle(model, n, S, dtype):
    dev = cuda()
    sd = init_vit_weights(model)
    g = torch.Generator().manual_seed(11)
    for k in sd:                     # non-trivial LayerNorm affine and biases: every folded term counts
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1 + 0.2 * torch.randn(sd[k].shape, generator=g)
    rng = np.random.default_rng(n)
    frame = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(40, 280, n), rng.uniform(40, 200, n),
                      rng.uniform(30, 90, n), rng.uniform(30, 90, n)], 1).astype(np.float32)
    crops = crop_resize(frame, boxes, size=S)
    want = torch.from_numpy(vit_forward(crops, sd, model))
    patches = torch.from_numpy(patchify(crops)).to(dtype).to(dev)
    encs = _encoders(model, n, S, sd, dev, {"plain": None, "kernel": "kernel", "epilogue": None}, dtype)
    assert (encs["plain"].ln_fold, encs["kernel"].ln_fold, encs["epilogue"].ln_fold) == (0, 1, 2)
    rel = {}
    for name, enc in encs.items():
        out = enc.forward(patches, n).cpu()
        rel[name] = _rel(out, want)
        cos = torch.nn.functional.cosine_similarity(out, want, dim=1).min()
        print(f"{model} S={S} {dtype} LayerNorm {name}: rel-L2 {rel[name]:.3e}, min cos {cos:.7f}")
        assert rel[name] < (1.5e-2 if dtype == torch.bfloat16 else 3e-3) and cos > 1 - 1e-3
    # folding removes the 16-bit rounding of the normalised activation: never further from the
    # oracle than the LayerNorm kernels by more than noise
    assert rel["kernel"] < rel["plain"] * 1.15 + 3e-4
    assert rel["epilogue"] < rel["plain"] * 1.15 + 3e-4
    assert encs["epilogue"].launches < encs["plain"].launches


def test_encoder_ln_fold_is_batch_invariant():
    """Sharding particles over GPUs relies on a crop's embedding not depending on its batch."""
    dev = cuda()
    model, n = "vit_tiny_patch16_224", 37
    sd = init_vit_weights(model)
    enc = _encoders(model, n, 224, sd, dev, {"epilogue": None})["epilogue"]
    g = torch.Generator().manual_seed(5)
    patches = torch.randn(n * 196, 768, generator=g).to(torch.bfloat16).to(dev)
    full = enc.forward(patches, n).clone()
    for lo, hi in ((0, 1), (5, 18), (30, 37)):
        part = enc.forward(patches[lo * 196:hi * 196].contiguous(), hi - lo)
        assert torch.equal(part, full[lo:hi])


def test_ln_fold_writers_stay_inside_their_buffers():
    """compute-sanitizer is closed on this pool: the new writers (row statistics, the residual
    epilogues' partial sums, the folded GEMM's TMA stores) get sentinel-filled arenas around their
    outputs, ragged row counts, and both epilogue flavours of the residual GEMM."""
    from vitparticlefiltertracker_b200 import _lib, ops
    dev = cuda()
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    PAD = 2048

    def arena(n_elems, dtype, fill):
        a = torch.full((n_elems + 2 * PAD,), fill, dtype=dtype, device=dev)
        return a, a[PAD:PAD + n_elems]

    def intact(a, n_elems, fill):
        return bool((a[:PAD] == fill).all()) and bool((a[PAD + n_elems:] == fill).all())

    g = torch.Generator().manual_seed(9)
    for M, D in ((1, 192), (129, 192), (333, 384), (257, 768), (700, 1024)):
        x = _stream_like(M, D, g, dev)
        # row statistics kernel
        a, st = arena(M * 2, torch.float32, -7.0)
        _lib.check(lib.vpf_rowstats16(x.data_ptr(), st.data_ptr(), M, D, stream), "vpf_rowstats16")
        torch.cuda.synchronize()
        assert intact(a, M * 2, -7.0) and bool((st.view(M, 2)[:, 1] > 0).all()), (M, D)
        # residual GEMM + partial sums: K = D (TMA epilogue) and K = 4 D (register-staged epilogue)
        P = ops.ln_stats_parts(D)
        for K in (D, 4 * D):
            act = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
            w = (torch.randn(D, K, generator=g) * 0.03).to(torch.bfloat16).to(dev)
            bias = torch.zeros(D, device=dev)
            ax, xs = arena(M * D, torch.float16, 77.0)
            xs.view(M, D).copy_(x)
            ap, parts = arena(P * M * 2, torch.float32, -3.0)
            _lib.check(lib.vpf_gemm_residual_stats(act.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                                   xs.data_ptr(), parts.data_ptr(), M, D, K,
                                                   _lib.DTYPE_BF16, stream), "vpf_gemm_residual_stats")
            torch.cuda.synchronize()
            assert intact(ax, M * D, 77.0) and intact(ap, P * M * 2, -3.0), (M, D, K)
            assert bool((parts.view(P, M, 2)[:, :, 1] >= 0).all()), (M, D, K)     # every slot written
            # folded GEMM reading those sums: ragged M, output fenced
            N = 256
            wq = (torch.randn(N, D, generator=g) * 0.05).to(torch.float16).to(dev)
            ao, out = arena(M * N, torch.bfloat16, 1000.0)
            _lib.check(lib.vpf_gemm_ln(xs.data_ptr(), wq.data_ptr(), torch.zeros(N, device=dev).data_ptr(),
                                       parts.data_ptr(), P, out.data_ptr(), M, N, D, 1,
                                       _lib.DTYPE_BF16, stream), "vpf_gemm_ln")
            torch.cuda.synchronize()
            assert intact(ao, M * N, 1000.0), (M, D, K)
            assert bool((out.float().abs() < 500.0).all()), (M, D, K)            # every element written


def test_ln_fold_argument_checks():
    from vitparticlefiltertracker_b200 import ops, pack_for_device
    from vitparticlefiltertracker_b200.vit_weights import pack_ln_fold
    dev = cuda()
    model = "vit_tiny_patch16_224"
    cfg = VIT_CONFIGS[model]
    sd = init_vit_weights(model)
    packed = pack_for_device(model, sd)
    with pytest.raises(ValueError):        # the fp32 stream cannot be a GEMM operand
        ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], 2, packed, dev, ln_fold=pack_ln_fold(model, sd))
    with pytest.raises(ValueError):        # wrong number of tensors
        ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], 2, packed, dev, residual_dtype=torch.float16,
                       ln_fold=pack_ln_fold(model, sd)[:-1])
    from vitparticlefiltertracker_b200 import Tracker
    with pytest.raises(ValueError):
        Tracker(model, 8, residual_dtype=torch.float32, ln_fold=True)
    assert Tracker(model, 8, ln_fold=False).encoder.ln_fold == 0
    assert Tracker(model, 8).encoder.ln_fold == 2
