"""timm-style ViT/16 forward in CPU PyTorch fp32 (oracle) -- test infrastructure only.

Restates SURVEY.md Appendix A D1/D2 (anchor: /root/reference/README.md:7, "Vision Transformer
(ViT): ... feature extraction"; no model code, `timm` pin or weights exist in the reference;
`timm` is not installed here, so its `vit_{tiny,base,large}_patch16_224` forward with
``num_classes=0, global_pool='token'`` is restated by hand and cross-checked against
torchvision's independent implementation in tests/test_oracle_vit.py).

state_dict key names are timm's:
  patch_embed.proj.{weight[D,3,16,16],bias}  cls_token[1,1,D]  pos_embed[1,197,D]
  blocks.{i}.norm1.{weight,bias}  blocks.{i}.attn.qkv.{weight[3D,D],bias}
  blocks.{i}.attn.proj.{weight,bias}  blocks.{i}.norm2.{weight,bias}
  blocks.{i}.mlp.fc1.{weight[4D,D],bias}  blocks.{i}.mlp.fc2.{weight[D,4D],bias}
  norm.{weight,bias}
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

VIT_CONFIGS = {
    "vit_tiny_patch16_224": dict(dim=192, depth=12, heads=3),
    "vit_small_patch16_224": dict(dim=384, depth=12, heads=6),
    "vit_base_patch16_224": dict(dim=768, depth=12, heads=12),
    "vit_large_patch16_224": dict(dim=1024, depth=24, heads=16),
}
TOKENS = 197
LN_EPS = 1e-6
WEIGHT_SEED = 1234


def vit_flops_per_crop(model: str, crop_size: int = 224) -> float:
    """2*P*768*D + L*(24*T*D^2 + 4*T^2*D), P = (S/16)^2 patches, T = P + 1 tokens (196 / 197 at
    S = 224): full forward, 2 FLOP/MAC, no padding."""
    c = VIT_CONFIGS[model]
    d, l = c["dim"], c["depth"]
    p = (crop_size // 16) ** 2
    t = p + 1
    return 2.0 * p * 768 * d + l * (24.0 * t * d * d + 4.0 * t * t * d)


def resample_pos_embed(pos_embed: torch.Tensor, grid: int) -> torch.Tensor:
    """pos_embed [1, 1+14*14, D] -> [1, 1+grid*grid, D] for a (16*grid)^2 input.

    Restates timm's ``resample_abs_pos_embed`` (what ``dynamic_img_size`` / a changed ``img_size``
    applies to `vit_*_patch16_224` weights): the class-token row is kept, the 14x14 grid part is
    resized with bicubic interpolation, antialias on, align_corners off, in fp32."""
    n_tok, d = pos_embed.shape[1], pos_embed.shape[2]
    old = int(round(math.sqrt(n_tok - 1)))
    assert old * old + 1 == n_tok, "pos_embed must be [1, 1+g*g, D]"
    if grid == old:
        return pos_embed
    cls_row, grid_rows = pos_embed[:, :1], pos_embed[:, 1:]
    g = grid_rows.reshape(1, old, old, d).permute(0, 3, 1, 2).to(torch.float32)
    g = F.interpolate(g, size=(grid, grid), mode="bicubic", antialias=True, align_corners=False)
    g = g.permute(0, 2, 3, 1).reshape(1, grid * grid, d)
    return torch.cat([cls_row.to(torch.float32), g], dim=1)


def weight_keys(model: str):
    """(key, shape, kind) in the fixed order the seeded initialiser walks them (D2)."""
    c = VIT_CONFIGS[model]
    d, l = c["dim"], c["depth"]
    keys = [
        ("patch_embed.proj.weight", (d, 3, 16, 16), "w"),
        ("patch_embed.proj.bias", (d,), "zero"),
        ("cls_token", (1, 1, d), "cls"),
        ("pos_embed", (1, TOKENS, d), "w"),
    ]
    for i in range(l):
        p = f"blocks.{i}."
        keys += [
            (p + "norm1.weight", (d,), "one"), (p + "norm1.bias", (d,), "zero"),
            (p + "attn.qkv.weight", (3 * d, d), "w"), (p + "attn.qkv.bias", (3 * d,), "zero"),
            (p + "attn.proj.weight", (d, d), "w"), (p + "attn.proj.bias", (d,), "zero"),
            (p + "norm2.weight", (d,), "one"), (p + "norm2.bias", (d,), "zero"),
            (p + "mlp.fc1.weight", (4 * d, d), "w"), (p + "mlp.fc1.bias", (4 * d,), "zero"),
            (p + "mlp.fc2.weight", (d, 4 * d), "w"), (p + "mlp.fc2.bias", (d,), "zero"),
        ]
    keys += [("norm.weight", (d,), "one"), ("norm.bias", (d,), "zero")]
    return keys


def init_vit_weights(model: str, seed: int = WEIGHT_SEED) -> dict:
    """Seeded random init (D2): weights/pos_embed trunc-normal(std .02, +-2 std), biases 0,
    LN gamma 1 / beta 0, cls_token normal(std 1e-6).  CPU generator => bit-reproducible."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for key, shape, kind in weight_keys(model):
        if kind == "w":
            t = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        elif kind == "cls":
            t = torch.empty(shape, dtype=torch.float32).normal_(0.0, 1e-6, generator=g)
        elif kind == "one":
            t = torch.ones(shape, dtype=torch.float32)
        else:
            t = torch.zeros(shape, dtype=torch.float32)
        sd[key] = t
    return sd


@torch.no_grad()
def vit_forward(crops, sd: dict, model: str, chunk: int = 32, return_tokens: bool = False):
    """crops fp32 [N,3,S,S] (numpy or torch; S = 224, or a smaller multiple of 16 with the
    position embedding resampled to the (S/16)^2 grid) -> CLS embedding fp32 [N,D] (numpy)."""
    c = VIT_CONFIGS[model]
    d, depth, heads = c["dim"], c["depth"], c["heads"]
    hd = d // heads
    x_all = torch.as_tensor(np.asarray(crops), dtype=torch.float32)
    grid = x_all.shape[-1] // 16
    TOKENS = grid * grid + 1                                   # noqa: N806 (197 at S = 224)
    pos_embed = resample_pos_embed(sd["pos_embed"], grid)
    outs = []
    for s in range(0, x_all.shape[0], chunk):
        x = x_all[s:s + chunk]
        n = x.shape[0]
        x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16)
        x = x.flatten(2).transpose(1, 2)                       # [n,P,D], patch index = py*grid+px
        x = torch.cat([sd["cls_token"].expand(n, -1, -1), x], dim=1) + pos_embed
        for i in range(depth):
            p = f"blocks.{i}."
            h = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
            qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
            qkv = qkv.reshape(n, TOKENS, 3, heads, hd).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]                   # [n,H,T,hd]
            att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
            att = att.softmax(dim=-1)
            o = (att @ v).transpose(1, 2).reshape(n, TOKENS, d)
            x = x + F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
            h = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
            h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
            h = F.gelu(h)                                      # exact (erf) GELU
            x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        if return_tokens:
            outs.append(x)
        else:
            x = F.layer_norm(x[:, 0], (d,), sd["norm.weight"], sd["norm.bias"], LN_EPS)
            outs.append(x)
    return torch.cat(outs, dim=0).numpy()
