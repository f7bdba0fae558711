"""CPU: timm-checkpoint handling -- the package's `load_checkpoint` and
the oracle's `normalise_state_dict` must accept what real checkpoints look like (wrapped,
prefixed, extra head / fc_norm keys, fp16 tensors) and agree with each other."""
import numpy as np
import pytest
import torch

from oracle.vit import init_vit_weights, normalise_state_dict, vit_forward
from checkpoint_util import realistic_checkpoint


def test_load_checkpoint_handles_real_world_layouts(tmp_path):
    from vitparticlefiltertracker_b200 import init_state_dict, load_checkpoint, pack_for_device
    model = "vit_tiny_patch16_224"
    ckpt, clean = realistic_checkpoint(model, seed=5)
    assert any(k.startswith("module.head.") for k in ckpt["state_dict"])
    assert ckpt["state_dict"]["module.blocks.0.attn.qkv.weight"].dtype == torch.float16
    path = tmp_path / "vit_tiny.pt"
    torch.save(ckpt, path)
    for src in (ckpt, str(path), path):
        sd = load_checkpoint(src, model)
        assert set(sd) == set(init_state_dict(model))            # extras dropped, nothing missing
        assert all(t.dtype == torch.float32 for t in sd.values())
        for k, t in sd.items():
            assert torch.equal(t, clean[k]), k
    ora = normalise_state_dict(ckpt, model)                       # the oracle reads the same file
    assert all(torch.equal(ora[k], sd[k]) for k in sd)
    packed = pack_for_device(model, sd)
    assert len(packed) == 4 + 12 * 12 + 2 and packed[0].dtype == torch.bfloat16


def test_load_checkpoint_errors_name_the_key():
    from vitparticlefiltertracker_b200 import load_checkpoint
    model = "vit_tiny_patch16_224"
    sd = init_vit_weights(model)
    bad = dict(sd); del bad["blocks.3.mlp.fc1.bias"]
    with pytest.raises(KeyError, match="blocks.3.mlp.fc1.bias"):
        load_checkpoint(bad, model)
    bad = dict(sd); bad["blocks.0.attn.qkv.weight"] = torch.zeros(10, 10)
    with pytest.raises(ValueError, match="blocks.0.attn.qkv.weight"):
        load_checkpoint(bad, model)
    with pytest.raises(ValueError, match="vit_base"):
        load_checkpoint(sd, "vit_base_patch16_224")                # a ViT-Ti file for a ViT-B model
    bad = dict(sd); bad["norm.weight"] = torch.full((192,), float("nan"))
    with pytest.raises(ValueError, match="non-finite"):
        load_checkpoint(bad, model)


def test_oracle_forward_uses_the_checkpoint_not_the_seeded_init():
    model = "vit_tiny_patch16_224"
    ckpt, _ = realistic_checkpoint(model, seed=5)
    x = np.random.default_rng(0).standard_normal((2, 3, 224, 224)).astype(np.float32)
    a = vit_forward(x, normalise_state_dict(ckpt, model), model)
    b = vit_forward(x, init_vit_weights(model), model)
    assert np.isfinite(a).all() and np.abs(a - b).max() > 1e-2
