"""A synthetic stand-in for a real timm ViT checkpoint (none exist offline): every tensor the
tracker reads is non-trivial (random biases, LayerNorm gains / shifts away from 1 / 0, a few
large-gain "outlier" channels of the kind trained ViTs have), stored in fp16, under a
``{"state_dict": {"module.<key>": ...}}`` wrapper, with the extra keys a classification
checkpoint carries (``head.*``, ``fc_norm.*``)."""
import torch

from oracle.vit import VIT_CONFIGS, init_vit_weights


def realistic_checkpoint(model: str, seed: int = 5):
    """Returns (checkpoint as saved, clean fp32 state_dict the loaders must produce)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = VIT_CONFIGS[model]["dim"]
    sd = init_vit_weights(model, seed=seed + 1000)
    for k, t in sd.items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            t[7] = 4.0                                       # an outlier channel
        elif k.endswith(".bias"):
            t.copy_(0.05 * torch.randn(t.shape, generator=g))
        elif k == "cls_token":
            t.copy_(0.02 * torch.randn(t.shape, generator=g))
    half = {k: t.to(torch.float16) for k, t in sd.items()}
    clean = {k: t.to(torch.float32) for k, t in half.items()}
    extra = {"head.weight": torch.randn(1000, d, generator=g).half(),
             "head.bias": torch.zeros(1000).half(),
             "fc_norm.weight": torch.ones(d).half(), "fc_norm.bias": torch.zeros(d).half()}
    wrapped = {"state_dict": {"module." + k: v for k, v in {**half, **extra}.items()},
               "epoch": 300, "arch": model}
    return wrapped, clean
