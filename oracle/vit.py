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


def normalise_state_dict(sd: dict, model: str) -> dict:
    """What a real timm checkpoint needs before the forward below can use it: unwrap a
    ``{"state_dict": ...}`` / ``{"model": ...}`` container, strip a ``module.`` prefix, drop the
    keys the tracker has no use for (``head.*``, ``fc_norm.*``, ...), check shapes, convert
    fp16 / bf16 tensors to fp32."""
    for k in ("state_dict", "model"):
        if k in sd and isinstance(sd[k], dict):
            sd = sd[k]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    out = {}
    for key, shape, _ in weight_keys(model):
        if key not in sd:
            raise KeyError(f"checkpoint has no {key!r} (expected timm {model} key names)")
        t = torch.as_tensor(sd[key]).detach().to(torch.float32)
        if key == "pos_embed":
            if t.dim() != 3 or t.shape[0] != 1 or t.shape[2] != shape[2]:
                raise ValueError(f"pos_embed has shape {tuple(t.shape)}")
        elif tuple(t.shape) != tuple(shape):
            raise ValueError(f"{key} has shape {tuple(t.shape)}, {model} needs {tuple(shape)}")
        out[key] = t.contiguous()
    return out


# ---- operand-precision emulation (study tool for SURVEY.md section 8f-2 "FP8 GEMMs") ----------
E4M3_MAX = 448.0


def quantise(t: torch.Tensor, fmt):
    """Rounds a GEMM operand to `fmt` and back to fp32: None (fp32), "bf16", "fp16", or "e4m3"
    (per-tensor scale amax -> 448, round-to-nearest-even to float8_e4m3fn, the recipe a
    `tcgen05.mma.kind::f8f6f4` path with per-tensor scaling would use)."""
    if fmt is None:
        return t
    if fmt == "bf16":
        return t.to(torch.bfloat16).to(torch.float32)
    if fmt == "fp16":
        return t.to(torch.float16).to(torch.float32)
    if fmt == "e4m3":
        amax = t.abs().max().clamp_min(1e-12)
        s = E4M3_MAX / amax
        return (t * s).to(torch.float8_e4m3fn).to(torch.float32) / s
    raise ValueError(f"unknown operand format {fmt!r}")


@torch.no_grad()
def vit_forward(crops, sd: dict, model: str, chunk: int = 32, return_tokens: bool = False,
                linear_fmt=None, attn_fmt=None):
    """crops fp32 [N,3,S,S] (numpy or torch; S = 224, or a smaller multiple of 16 with the
    position embedding resampled to the (S/16)^2 grid) -> CLS embedding fp32 [N,D] (numpy).

    `linear_fmt` / `attn_fmt` (default None = the fp32 oracle) emulate reduced-precision GEMM
    operands with fp32 accumulation -- the patch-embed / qkv / proj / fc1 / fc2 contractions and
    the two attention contractions respectively (see `quantise`); a CPU study tool
    (tools/fp8_study.py), not a code path of the tracker."""
    c = VIT_CONFIGS[model]
    d, depth, heads = c["dim"], c["depth"], c["heads"]
    hd = d // heads
    x_all = torch.as_tensor(np.asarray(crops), dtype=torch.float32)
    grid = x_all.shape[-1] // 16
    TOKENS = grid * grid + 1                                   # noqa: N806 (197 at S = 224)
    pos_embed = resample_pos_embed(sd["pos_embed"], grid)
    qw = {}

    def lin(h, wkey, bkey):
        if linear_fmt is None:
            return F.linear(h, sd[wkey], sd[bkey])
        if wkey not in qw:
            qw[wkey] = quantise(sd[wkey], linear_fmt)
        return F.linear(quantise(h, linear_fmt), qw[wkey], sd[bkey])

    outs = []
    for s in range(0, x_all.shape[0], chunk):
        x = x_all[s:s + chunk]
        n = x.shape[0]
        if linear_fmt is None:
            x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16)
        else:      # crops reach the GPU GEMM as 16-bit patches: keep them at least that coarse
            x = F.conv2d(quantise(x, "bf16" if linear_fmt == "e4m3" else linear_fmt),
                         quantise(sd["patch_embed.proj.weight"], "bf16" if linear_fmt == "e4m3"
                                  else linear_fmt), sd["patch_embed.proj.bias"], stride=16)
        x = x.flatten(2).transpose(1, 2)                       # [n,P,D], patch index = py*grid+px
        x = torch.cat([sd["cls_token"].expand(n, -1, -1), x], dim=1) + pos_embed
        for i in range(depth):
            p = f"blocks.{i}."
            h = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
            qkv = lin(h, p + "attn.qkv.weight", p + "attn.qkv.bias")
            qkv = quantise(qkv, attn_fmt)
            qkv = qkv.reshape(n, TOKENS, 3, heads, hd).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]                   # [n,H,T,hd]
            att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
            att = att.softmax(dim=-1)
            o = (quantise(att, attn_fmt) @ v).transpose(1, 2).reshape(n, TOKENS, d)
            x = x + lin(o, p + "attn.proj.weight", p + "attn.proj.bias")
            h = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
            h = lin(h, p + "mlp.fc1.weight", p + "mlp.fc1.bias")
            h = F.gelu(h)                                      # exact (erf) GELU
            x = x + lin(h, p + "mlp.fc2.weight", p + "mlp.fc2.bias")
        if return_tokens:
            outs.append(x)
        else:
            x = F.layer_norm(x[:, 0], (d,), sd["norm.weight"], sd["norm.bias"], LN_EPS)
            outs.append(x)
    return torch.cat(outs, dim=0).numpy()
