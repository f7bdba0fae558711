"""ViT/16 weight handling for the tracker: model table, seeded init
and packing of a timm-keyed state_dict into the pointer order libvpf expects (include/vpf.h).

No pretrained checkpoints exist offline and the upstream repository ships none
(/root/reference/README.md:7 names the model family only), so the default is a seeded random
init; ``state_dict=`` accepts real timm weights (`vit_*_patch16_224`, num_classes=0).
"""
from __future__ import annotations

import torch

VIT_CONFIGS = {
    "vit_tiny_patch16_224": dict(dim=192, depth=12, heads=3),
    "vit_small_patch16_224": dict(dim=384, depth=12, heads=6),
    "vit_base_patch16_224": dict(dim=768, depth=12, heads=12),
    "vit_large_patch16_224": dict(dim=1024, depth=24, heads=16),
}
TOKENS = 197
WEIGHT_SEED = 1234


def flops_per_crop(model: str, crop_size: int = 224) -> float:
    """Full forward, 2 FLOP/MAC, no padding: 2*P*768*D + L*(24*T*D^2 + 4*T^2*D) with
    P = (crop_size/16)^2 patches and T = P + 1 tokens (196 / 197 at the default 224)."""
    c = VIT_CONFIGS[model]
    d, l = c["dim"], c["depth"]
    p = (crop_size // 16) ** 2
    t = p + 1
    return 2.0 * p * 768 * d + l * (24.0 * t * d * d + 4.0 * t * t * d)


def resample_pos_embed(pos_embed: torch.Tensor, grid: int) -> torch.Tensor:
    """[1, 1+g0*g0, D] -> [1, 1+grid*grid, D]: what timm applies to `vit_*_patch16_224` weights when
    the input size changes (class-token row kept; grid rows resized bicubic, antialias, fp32).
    One-time host-side weight preprocessing, outside the per-frame path."""
    n_tok, d = pos_embed.shape[1], pos_embed.shape[2]
    g0 = int(round((n_tok - 1) ** 0.5))
    if g0 * g0 + 1 != n_tok:
        raise ValueError("pos_embed must be [1, 1+g*g, D]")
    if grid == g0:
        return pos_embed
    pe = pos_embed.detach().to(torch.float32).cpu()
    g = pe[:, 1:].reshape(1, g0, g0, d).permute(0, 3, 1, 2)
    g = torch.nn.functional.interpolate(g, size=(grid, grid), mode="bicubic", antialias=True,
                                        align_corners=False)
    return torch.cat([pe[:, :1], g.permute(0, 2, 3, 1).reshape(1, grid * grid, d)], dim=1)


def _block_keys(i):
    p = f"blocks.{i}."
    return [p + "norm1.weight", p + "norm1.bias", p + "attn.qkv.weight", p + "attn.qkv.bias",
            p + "attn.proj.weight", p + "attn.proj.bias", p + "norm2.weight", p + "norm2.bias",
            p + "mlp.fc1.weight", p + "mlp.fc1.bias", p + "mlp.fc2.weight", p + "mlp.fc2.bias"]


def init_state_dict(model: str, seed: int = WEIGHT_SEED) -> dict:
    """Weights / pos_embed trunc-normal(std .02, cut at +-.04), biases 0, LN (1, 0),
    cls_token normal(std 1e-6); one CPU generator walked in timm key order."""
    c = VIT_CONFIGS[model]
    d, depth = c["dim"], c["depth"]
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def tn(*shape):
        t = torch.empty(shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        return t

    sd = {"patch_embed.proj.weight": tn(d, 3, 16, 16),
          "patch_embed.proj.bias": torch.zeros(d),
          "cls_token": torch.empty(1, 1, d).normal_(0.0, 1e-6, generator=g),
          "pos_embed": tn(1, TOKENS, d)}
    for i in range(depth):
        k = _block_keys(i)
        sd[k[0]] = torch.ones(d); sd[k[1]] = torch.zeros(d)
        sd[k[2]] = tn(3 * d, d); sd[k[3]] = torch.zeros(3 * d)
        sd[k[4]] = tn(d, d); sd[k[5]] = torch.zeros(d)
        sd[k[6]] = torch.ones(d); sd[k[7]] = torch.zeros(d)
        sd[k[8]] = tn(4 * d, d); sd[k[9]] = torch.zeros(4 * d)
        sd[k[10]] = tn(d, 4 * d); sd[k[11]] = torch.zeros(d)
    sd["norm.weight"] = torch.ones(d)
    sd["norm.bias"] = torch.zeros(d)
    return sd


def _expected_shapes(model: str) -> dict:
    c = VIT_CONFIGS[model]
    d = c["dim"]
    shapes = {"patch_embed.proj.weight": (d, 3, 16, 16), "patch_embed.proj.bias": (d,),
              "cls_token": (1, 1, d), "pos_embed": None, "norm.weight": (d,), "norm.bias": (d,)}
    per_block = [(d,), (d,), (3 * d, d), (3 * d,), (d, d), (d,), (d,), (d,), (4 * d, d), (4 * d,),
                 (d, 4 * d), (d,)]
    for i in range(c["depth"]):
        shapes.update(zip(_block_keys(i), per_block))
    return shapes


def load_checkpoint(src, model: str) -> dict:
    """A timm `vit_*_patch16_224` checkpoint -> the fp32 state_dict the tracker packs.

    `src`: a path for `torch.load`, or a dict.  Real checkpoints come wrapped
    (``{"state_dict": ...}`` / ``{"model": ...}``), prefixed (``module.``), with keys the tracker
    has no use for (``head.*``, ``fc_norm.*``: the likelihood reads the CLS embedding after the
    final norm, i.e. `num_classes=0, global_pool='token'`) and often in fp16 / bf16: all handled
    here.  A missing key or a wrong shape raises with the key's name."""
    sd = torch.load(src, map_location="cpu", weights_only=True) if isinstance(src, (str, bytes)) \
        or hasattr(src, "__fspath__") else src
    for k in ("state_dict", "model"):
        if k in sd and isinstance(sd[k], dict):
            sd = sd[k]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    out = {}
    for key, shape in _expected_shapes(model).items():
        if key not in sd:
            raise KeyError(f"checkpoint has no {key!r} (expected timm {model} key names)")
        t = torch.as_tensor(sd[key]).detach().to(torch.float32)
        if key == "pos_embed":
            g = int(round((t.shape[1] - 1) ** 0.5)) if t.dim() == 3 else 0
            if t.dim() != 3 or t.shape[0] != 1 or t.shape[2] != VIT_CONFIGS[model]["dim"] \
                    or g * g + 1 != t.shape[1]:
                raise ValueError(f"pos_embed has shape {tuple(t.shape)}, need [1, 1+g*g, D]")
        elif tuple(t.shape) != shape:
            raise ValueError(f"{key} has shape {tuple(t.shape)}, {model} needs {shape}")
        if not torch.isfinite(t).all():
            raise ValueError(f"{key} holds non-finite values")
        out[key] = t.contiguous()
    return out


def pack_for_device(model: str, sd: dict, dtype=torch.bfloat16, crop_size: int = 224) -> list:
    """timm state_dict -> list of CPU tensors in libvpf pointer order: GEMM weights in the 16-bit
    operand type (`dtype`: bf16 or fp16) [out][in] (the K-major B operand), everything else fp32.
    `crop_size` < 224 resamples pos_embed to the (crop_size/16)^2 grid."""
    c = VIT_CONFIGS[model]
    d, depth = c["dim"], c["depth"]
    pos = resample_pos_embed(sd["pos_embed"], crop_size // 16)
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    bf = lambda t: t.detach().to(torch.float32).to(dtype).contiguous()
    out = [bf(sd["patch_embed.proj.weight"].reshape(d, 768)), f32(sd["patch_embed.proj.bias"]),
           f32(sd["cls_token"].reshape(d)), f32(pos.reshape(pos.shape[1], d))]
    for i in range(depth):
        for key in _block_keys(i):
            t = sd[key]
            out.append(bf(t) if (key.endswith("weight") and t.dim() == 2) else f32(t))
    out += [f32(sd["norm.weight"]), f32(sd["norm.bias"])]
    return out


def pack_ln_fold(model: str, sd: dict, dtype=torch.bfloat16, storage=torch.float16) -> list:
    """LayerNorm folded into the two GEMMs that consume it (include/vpf.h `vpf_vit_set_ln_fold`):

        LN(x) W^T + b = r (x W'^T) + b',   W'_jk = gamma_k W_jk - mean_k(gamma_k W_jk),   b' = b + W beta

    with r the row's 1/sqrt(var + eps): subtracting each weight row's mean over k makes x W'^T equal
    (x - mean(x)) (gamma (.) W)^T, so the GEMM's epilogue needs only r.  Per block four CPU tensors:
    qkv W', b' (fp32), fc1 W', b'.  Formed in fp64 from the fp32 master weights.

    W' is rounded to the encoder's operand type `dtype` and then STORED in fp16 (exact for bf16
    values above fp16's subnormal range): the folded GEMMs read the fp16 residual stream as A, and
    tcgen05's kind::f16 wants one 16-bit type for both operands (A = f16 with B = bf16 faults on
    sm_100a).  Keeping bf16's 8 significant bits inside the fp16 container is deliberate: the step
    is power-limited and the multipliers' switching grows with the operands' mantissa width --
    measured on one B200, alternating: 146.5 / 147.2 ms per step against 150.8 / 150.8 ms with full
    fp16 weights (155.3 / 155.4 ms with separate LayerNorm kernels).

    The rounding leaves a row sum eps_j = sum_k W'_jk of ~3e-2 |w| (bf16; 4e-3 |w| fp16) instead of
    0, i.e. an output error of r mean(x) eps_j ~ 1e-3 (mean(x) / std(x)) relative to the output's
    scale -- a fraction of the 16-bit output rounding for the |mean| < std rows of a ViT stream.
    (`storage=torch.float32, dtype=torch.float32` gives the unrounded algebra for host tests.)"""
    out = []
    for i in range(VIT_CONFIGS[model]["depth"]):
        k = _block_keys(i)
        for g_key, b_key, w_key, bias_key in ((k[0], k[1], k[2], k[3]), (k[6], k[7], k[8], k[9])):
            gamma, beta = sd[g_key].detach().double(), sd[b_key].detach().double()
            w, bias = sd[w_key].detach().double(), sd[bias_key].detach().double()
            wf = w * gamma[None, :]
            wf = (wf - wf.mean(dim=1, keepdim=True)).to(torch.float32).to(dtype).to(torch.float32)
            out += [wf.to(storage).contiguous(), (bias + w @ beta).to(torch.float32).contiguous()]
    return out
