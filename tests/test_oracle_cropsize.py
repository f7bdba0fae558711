"""The oracle's variable crop size ("native-resolution crops"): the same
sampling rule on an S x S grid, (S/16)^2 patches, position embedding resampled the way timm does
for `vit_*_patch16_224` weights at another input size.  Pins: an independent grid_sample
formulation, torchvision's VisionTransformer at image_size = S, analytic identities, and the
committed golden vectors tests/golden/oracle_v2_crop.npz (tests/golden/make_golden_crop.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.crop import crop_resize, crop_resize_grid_sample, patchify
from oracle.tracker import OracleTracker
from oracle.vit import (VIT_CONFIGS, init_vit_weights, resample_pos_embed, vit_flops_per_crop,
                        vit_forward)

MODEL = "vit_tiny_patch16_224"


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "oracle_v2_crop.npz"))


@pytest.mark.parametrize("S", [16, 64, 112, 208])
def test_crop_size_matches_grid_sample(S):
    f = np.random.default_rng(S).integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
    boxes = np.array([[64, 48, 40, 30], [10, 10, 50, 50], [120, 90, 64, 64], [30.25, 60.5, 17.3, 9.1]],
                     dtype=np.float32)
    a = crop_resize(f, boxes, size=S)
    assert a.shape == (4, 3, S, S) and a.dtype == np.float32
    assert np.abs(a - crop_resize_grid_sample(f, boxes, size=S)).max() < 2e-3
    assert patchify(a).shape == (4 * (S // 16) ** 2, 768)


def test_native_resolution_crop_is_the_source_pixels():
    """A 64x64 box at integer centre sampled on a 64x64 grid reproduces the frame window."""
    f = np.random.default_rng(1).integers(0, 256, size=(100, 120, 3), dtype=np.uint8)
    out = crop_resize(f, np.array([[60, 50, 64, 64]], dtype=np.float32), size=64)[0]
    want = np.transpose(f[18:82, 28:92].astype(np.float32), (2, 0, 1)) * np.float32(2 / 255) - 1
    assert np.abs(out - want).max() < 1e-5


@pytest.mark.parametrize("bad", [0, 8, 100, 240])
def test_bad_crop_size_raises(bad):
    with pytest.raises(ValueError):
        crop_resize(np.zeros((32, 32, 3), np.uint8), np.array([[16, 16, 8, 8]], np.float32), size=bad)


def test_resample_pos_embed_properties():
    pe = init_vit_weights(MODEL)["pos_embed"]
    assert resample_pos_embed(pe, 14) is pe                       # native grid: untouched
    r = resample_pos_embed(pe, 4)
    assert r.shape == (1, 17, 192) and r.dtype == torch.float32
    assert torch.equal(r[:, 0], pe[:, 0])                          # class-token row kept
    # a constant grid stays constant; a linear ramp keeps its mean (antialiased bicubic)
    const = torch.cat([pe[:, :1], torch.full((1, 196, 192), 0.25)], dim=1)
    assert torch.allclose(resample_pos_embed(const, 7)[:, 1:], torch.tensor(0.25), atol=1e-6)
    ramp = torch.arange(14.0).repeat(14, 1).reshape(1, 196, 1).expand(1, 196, 192)
    rr = resample_pos_embed(torch.cat([pe[:, :1], ramp], dim=1), 7)[0, 1:, 0].reshape(7, 7)
    assert abs(float(rr.mean()) - 6.5) < 1e-4 and (rr[:, 1:] > rr[:, :-1]).all()


def test_vit_small_input_matches_torchvision():
    """torchvision's VisionTransformer built at image_size = 64 with the oracle's weights and the
    oracle's resampled pos_embed must give the oracle's 17-token embedding."""
    from torchvision.models.vision_transformer import VisionTransformer
    cfg = VIT_CONFIGS[MODEL]
    sd = init_vit_weights(MODEL)
    g = torch.Generator().manual_seed(0)
    for k in sd:                                                   # non-trivial biases / gains
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    d = cfg["dim"]
    tv = VisionTransformer(image_size=64, patch_size=16, num_layers=cfg["depth"],
                           num_heads=cfg["heads"], hidden_dim=d, mlp_dim=4 * d)
    tv.heads = torch.nn.Identity()
    m = {"conv_proj.weight": sd["patch_embed.proj.weight"], "conv_proj.bias": sd["patch_embed.proj.bias"],
         "class_token": sd["cls_token"], "encoder.pos_embedding": resample_pos_embed(sd["pos_embed"], 4),
         "encoder.ln.weight": sd["norm.weight"], "encoder.ln.bias": sd["norm.bias"]}
    for i in range(cfg["depth"]):
        p, q = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        m.update({q + "ln_1.weight": sd[p + "norm1.weight"], q + "ln_1.bias": sd[p + "norm1.bias"],
                  q + "self_attention.in_proj_weight": sd[p + "attn.qkv.weight"],
                  q + "self_attention.in_proj_bias": sd[p + "attn.qkv.bias"],
                  q + "self_attention.out_proj.weight": sd[p + "attn.proj.weight"],
                  q + "self_attention.out_proj.bias": sd[p + "attn.proj.bias"],
                  q + "ln_2.weight": sd[p + "norm2.weight"], q + "ln_2.bias": sd[p + "norm2.bias"],
                  q + "mlp.0.weight": sd[p + "mlp.fc1.weight"], q + "mlp.0.bias": sd[p + "mlp.fc1.bias"],
                  q + "mlp.3.weight": sd[p + "mlp.fc2.weight"], q + "mlp.3.bias": sd[p + "mlp.fc2.bias"]})
    tv.load_state_dict(m)
    tv.eval()
    x = torch.randn(3, 3, 64, 64, generator=g).clamp(-1, 1)
    with torch.no_grad():
        want = tv(x).numpy()
    got = vit_forward(x.numpy(), sd, MODEL)
    assert got.shape == (3, d)
    assert np.abs(got - want).max() < 2e-4


def test_flops_formula_by_crop_size():
    assert vit_flops_per_crop("vit_base_patch16_224", 224) == vit_flops_per_crop("vit_base_patch16_224")
    d, l, p, t = 768, 12, 16, 17
    assert vit_flops_per_crop("vit_base_patch16_224", 64) == 2.0 * p * 768 * d + l * (24.0 * t * d * d + 4.0 * t * t * d)


def test_crop_size_golden(G):
    sd = init_vit_weights(MODEL)
    for S in (64, 112):
        crops = crop_resize(G["crop_frame"], G["crop_boxes"], size=S)
        assert np.allclose(crops[:, :, ::5, ::7], G[f"crop_samples_{S}"], atol=1e-6)
        cs = np.array([crops.astype(np.float64).sum(), (crops.astype(np.float64) ** 2).sum()])
        assert np.allclose(cs, G[f"crop_checksum_{S}"], rtol=1e-9)
        assert np.allclose(patchify(crops)[::3, ::41], G[f"patch_samples_{S}"], atol=1e-6)
        assert np.allclose(resample_pos_embed(sd["pos_embed"], S // 16)[0, :, :8].numpy(),
                           G[f"pos_embed_{S}"], atol=1e-6)
        assert np.allclose(vit_forward(crops, sd, MODEL), G[f"vit_emb_{S}"], atol=2e-4)


def test_tracker_crop64_golden(G):
    from vitparticlefiltertracker_b200.synthetic import make_clip
    frames, gt = make_clip("box", 224, 224, 4, seed=0, target=(48, 48))
    trk = OracleTracker(MODEL, 16, seed=0, crop_size=64)
    trk.init(frames[0], gt[0])
    assert np.allclose(trk.template, G["track_template_64"], atol=2e-4)
    traj = []
    for t in range(1, 4):
        s = trk.step(frames[t])
        traj.append([*s.bbox, s.ess])
    assert np.allclose(np.array(traj), G["track_traj_64"], rtol=2e-3, atol=2e-3)


def test_product_pos_embed_resampling_matches_oracle():
    """The product's one-time weight preprocessing (vit_weights.resample_pos_embed / pack_for_device)
    is checked against the oracle's restatement; the product never imports oracle/."""
    from vitparticlefiltertracker_b200.vit_weights import (flops_per_crop, pack_for_device,
                                                           resample_pos_embed as product_resample)
    sd = init_vit_weights(MODEL)
    for grid in (1, 4, 7, 13, 14):
        assert torch.equal(product_resample(sd["pos_embed"], grid), resample_pos_embed(sd["pos_embed"], grid))
    packed = pack_for_device(MODEL, sd, crop_size=64)
    assert tuple(packed[3].shape) == (17, 192)
    assert torch.equal(packed[3], resample_pos_embed(sd["pos_embed"], 4)[0])
    assert tuple(pack_for_device(MODEL, sd)[3].shape) == (197, 192)
    for S in (64, 224):
        assert flops_per_crop("vit_base_patch16_224", S) == vit_flops_per_crop("vit_base_patch16_224", S)
