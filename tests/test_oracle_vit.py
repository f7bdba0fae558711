"""Cross-checks the hand-written timm-style oracle ViT against torchvision's independent one."""
import numpy as np
import pytest
import torch

from oracle.vit import VIT_CONFIGS, init_vit_weights, vit_forward, vit_flops_per_crop, weight_keys


def _torchvision_vit(model, sd):
    from torchvision.models.vision_transformer import VisionTransformer
    c = VIT_CONFIGS[model]
    tv = VisionTransformer(image_size=224, patch_size=16, num_layers=c["depth"],
                           num_heads=c["heads"], hidden_dim=c["dim"], mlp_dim=4 * c["dim"],
                           num_classes=1000)
    tv.heads = torch.nn.Identity()
    m = {"conv_proj.weight": sd["patch_embed.proj.weight"],
         "conv_proj.bias": sd["patch_embed.proj.bias"],
         "class_token": sd["cls_token"], "encoder.pos_embedding": sd["pos_embed"],
         "encoder.ln.weight": sd["norm.weight"], "encoder.ln.bias": sd["norm.bias"]}
    for i in range(c["depth"]):
        s, t = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        m[t + "ln_1.weight"] = sd[s + "norm1.weight"]; m[t + "ln_1.bias"] = sd[s + "norm1.bias"]
        m[t + "self_attention.in_proj_weight"] = sd[s + "attn.qkv.weight"]
        m[t + "self_attention.in_proj_bias"] = sd[s + "attn.qkv.bias"]
        m[t + "self_attention.out_proj.weight"] = sd[s + "attn.proj.weight"]
        m[t + "self_attention.out_proj.bias"] = sd[s + "attn.proj.bias"]
        m[t + "ln_2.weight"] = sd[s + "norm2.weight"]; m[t + "ln_2.bias"] = sd[s + "norm2.bias"]
        m[t + "mlp.0.weight"] = sd[s + "mlp.fc1.weight"]; m[t + "mlp.0.bias"] = sd[s + "mlp.fc1.bias"]
        m[t + "mlp.3.weight"] = sd[s + "mlp.fc2.weight"]; m[t + "mlp.3.bias"] = sd[s + "mlp.fc2.bias"]
    missing, unexpected = tv.load_state_dict(m, strict=False)
    assert not missing, missing
    return tv.eval()


def _randomise(sd, seed=3):
    """Non-trivial biases / LN params so the comparison exercises every term."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith("bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.05
        elif "norm" in k and k.endswith("weight"):
            out[k] = 1 + torch.randn(v.shape, generator=g) * 0.1
        elif k == "cls_token":
            out[k] = torch.randn(v.shape, generator=g) * 0.02
        else:
            out[k] = v * 3.0
    return out


def test_oracle_vit_matches_torchvision_tiny():
    model = "vit_tiny_patch16_224"
    sd = _randomise(init_vit_weights(model))
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    want = _torchvision_vit(model, sd)(x).detach().numpy()
    got = vit_forward(x.numpy(), sd, model)
    assert got.shape == (3, 192)
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()


def test_init_is_deterministic_and_matches_spec():
    a = init_vit_weights("vit_tiny_patch16_224")
    b = init_vit_weights("vit_tiny_patch16_224")
    assert list(a) == [k for k, _, _ in weight_keys("vit_tiny_patch16_224")]
    for k in a:
        assert torch.equal(a[k], b[k])
    w = a["blocks.0.attn.qkv.weight"]
    assert w.abs().max() <= 0.04 and 0.015 < w.std() < 0.02
    assert a["blocks.0.attn.qkv.bias"].abs().max() == 0
    assert torch.equal(a["norm.weight"], torch.ones(192))
    assert a["cls_token"].abs().max() < 1e-5


def test_flops_formula():
    assert abs(vit_flops_per_crop("vit_tiny_patch16_224") / 1e9 - 2.507) < 0.01
    assert abs(vit_flops_per_crop("vit_base_patch16_224") / 1e9 - 35.126) < 0.01
    assert abs(vit_flops_per_crop("vit_large_patch16_224") / 1e9 - 123.107) < 0.01


def test_operand_quantisation_emulation():
    """oracle.vit.quantise (the FP8 study's emulation): e4m3 keeps 3
    mantissa bits whatever the per-tensor scale; bf16 / fp16 match torch's own conversions."""
    import torch
    from oracle.vit import quantise
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, generator=g) * 3.7
    assert quantise(x, None) is x
    assert torch.equal(quantise(x, "bf16"), x.to(torch.bfloat16).float())
    assert torch.equal(quantise(x, "fp16"), x.to(torch.float16).float())
    q = quantise(x, "e4m3")
    big = x.abs() > x.abs().max() * 2.0 ** -6           # the normal range of the scaled format
    rel = ((q - x).abs() / x.abs())[big]
    assert float(rel.max()) <= 2.0 ** -4 + 1e-6 and float(rel.mean()) > 2.0 ** -7
    assert float(quantise(x * 1000, "e4m3").abs().max()) == pytest.approx(float((x * 1000).abs().max()), rel=1e-6)
    with pytest.raises(ValueError):
        quantise(x, "int4")
