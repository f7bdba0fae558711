"""Host-side check of the LayerNorm folding algebra (vit_weights.pack_ln_fold) against the oracle's
own norm -> linear (oracle/vit.py), in fp64: no GPU needed."""
import torch

from oracle.vit import init_vit_weights
from vitparticlefiltertracker_b200.vit_weights import VIT_CONFIGS, pack_ln_fold


def test_pack_ln_fold_reproduces_layernorm_then_linear():
    model = "vit_tiny_patch16_224"
    d, depth = VIT_CONFIGS[model]["dim"], VIT_CONFIGS[model]["depth"]
    sd = init_vit_weights(model)
    g = torch.Generator().manual_seed(3)
    for k in sd:
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1 + 0.3 * torch.randn(sd[k].shape, generator=g)
    fold = pack_ln_fold(model, sd, torch.float32, torch.float32)   # unrounded W': the algebra alone
    assert len(fold) == 4 * depth
    x = (torch.randn(50, d, generator=g) * 3 + 0.8).double()
    r = 1.0 / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + 1e-6)
    for blk in (0, depth - 1):
        for j, (nk, lk) in enumerate((("norm1", "attn.qkv"), ("norm2", "mlp.fc1"))):
            p = f"blocks.{blk}."
            ref = torch.nn.functional.layer_norm(x, (d,), sd[p + nk + ".weight"].double(),
                                                 sd[p + nk + ".bias"].double(), 1e-6) \
                @ sd[p + lk + ".weight"].double().t() + sd[p + lk + ".bias"].double()
            wf, bf = (t.double() for t in fold[blk * 4 + j * 2: blk * 4 + j * 2 + 2])
            got = r * (x @ wf.t()) + bf[None]          # the row mean is never needed
            assert (got - ref).abs().max() < 1e-5


def test_pack_ln_fold_layout_and_centring():
    model = "vit_tiny_patch16_224"
    sd = init_vit_weights(model)
    fold = pack_ln_fold(model, sd, torch.float16)
    assert fold[0].dtype == torch.float16 and fold[1].dtype == torch.float32
    rounded = pack_ln_fold(model, sd)                  # bf16 precision in the fp16 container
    assert rounded[0].dtype == torch.float16
    assert torch.equal(rounded[0].float(), rounded[0].float().to(torch.bfloat16).float())
    assert (rounded[0].float() - fold[0].float()).abs().max() <= 0.08 * 2.0 ** -9
    assert fold[0].shape == (3 * 192, 192) and fold[2].shape == (4 * 192, 192)
    assert fold[1].shape == (3 * 192,) and fold[3].shape == (4 * 192,)
    # rows sum to zero up to the fp16 rounding of 192 weights of magnitude <= 0.04
    assert fold[0].double().sum(1).abs().max() < 192 * 0.04 * 2.0 ** -11
