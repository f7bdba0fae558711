"""Library comparator (informative): the same ViT-B/16 encoder forward on
N crops written with stock PyTorch bf16 ops (cuBLAS GEMMs via F.linear, fused SDPA attention,
F.layer_norm, F.gelu; fp32 residual stream like ours) against libvpf's encoder, same weights, same
patches, same GPU, alternating.  Test-only code: the product never calls any of this."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import VIT_CONFIGS, flops_per_crop, init_state_dict, ops, pack_for_device  # noqa: E402


def torch_forward(patches, sd, cfg, n, chunk):
    """patches bf16 [n*196, 768] -> CLS embedding fp32 [n, D]; weights bf16, residual fp32."""
    d, depth, heads = cfg["dim"], cfg["depth"], cfg["heads"]
    outs = []
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        p = patches[s * 196:(s + m) * 196]
        x = F.linear(p, sd["patch_w"], None).float().view(m, 196, d) + sd["patch_b"]
        x = torch.cat([sd["cls"].expand(m, 1, d), x], dim=1) + sd["pos"]
        for i in range(depth):
            w = sd["blocks"][i]
            h = F.layer_norm(x, (d,), w["ln1_g"], w["ln1_b"], 1e-6).to(torch.bfloat16)
            qkv = F.linear(h, w["qkv_w"], w["qkv_b"]).view(m, 197, 3, heads, 64).permute(2, 0, 3, 1, 4)
            o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            o = o.transpose(1, 2).reshape(m, 197, d)
            x = x + F.linear(o, w["proj_w"], w["proj_b"]).float()
            h = F.layer_norm(x, (d,), w["ln2_g"], w["ln2_b"], 1e-6).to(torch.bfloat16)
            h = F.gelu(F.linear(h, w["fc1_w"], w["fc1_b"]))
            x = x + F.linear(h, w["fc2_w"], w["fc2_b"]).float()
        outs.append(F.layer_norm(x[:, 0], (d,), sd["norm_g"], sd["norm_b"], 1e-6))
    return torch.cat(outs)


def main():
    model = "vit_base_patch16_224"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    dev = torch.device("cuda")
    cfg = VIT_CONFIGS[model]
    raw = init_state_dict(model)
    bf = lambda t: t.to(dev).to(torch.bfloat16)     # noqa: E731
    f32 = lambda t: t.to(dev).float()               # noqa: E731
    sd = {"patch_w": bf(raw["patch_embed.proj.weight"].reshape(cfg["dim"], 768)),
          "patch_b": f32(raw["patch_embed.proj.bias"]), "cls": f32(raw["cls_token"]),
          "pos": f32(raw["pos_embed"]), "norm_g": f32(raw["norm.weight"]), "norm_b": f32(raw["norm.bias"]),
          "blocks": []}
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        sd["blocks"].append({
            "ln1_g": f32(raw[p + "norm1.weight"]), "ln1_b": f32(raw[p + "norm1.bias"]),
            "qkv_w": bf(raw[p + "attn.qkv.weight"]), "qkv_b": bf(raw[p + "attn.qkv.bias"]),
            "proj_w": bf(raw[p + "attn.proj.weight"]), "proj_b": bf(raw[p + "attn.proj.bias"]),
            "ln2_g": f32(raw[p + "norm2.weight"]), "ln2_b": f32(raw[p + "norm2.bias"]),
            "fc1_w": bf(raw[p + "mlp.fc1.weight"]), "fc1_b": bf(raw[p + "mlp.fc1.bias"]),
            "fc2_w": bf(raw[p + "mlp.fc2.weight"]), "fc2_b": bf(raw[p + "mlp.fc2.bias"])})
    enc = ops.VitEncoder(cfg["dim"], cfg["depth"], cfg["heads"], n, pack_for_device(model, raw), dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    patches = (torch.randn(n * 196, 768, generator=g) * 0.5).clamp(-1, 1).to(torch.bfloat16).to(dev)
    emb = torch.empty(n, cfg["dim"], device=dev)

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    with torch.no_grad():
        ref = torch_forward(patches, sd, cfg, n, chunk)
        ours = enc.forward(patches, n, emb).clone()
        cos = F.cosine_similarity(ours, ref, dim=1)
        res = {"model": model, "n_crops": n, "torch_chunk": chunk,
               "agreement": {"rel_l2": float((ours - ref).norm() / ref.norm()), "min_cos": float(cos.min())},
               "ms": {"torch_eager_bf16": [], "libvpf": []}}
        for _ in range(3):
            res["ms"]["torch_eager_bf16"].append(round(timed(lambda: torch_forward(patches, sd, cfg, n, chunk), 3), 2))
            res["ms"]["libvpf"].append(round(timed(lambda: enc.forward(patches, n, emb), 3), 2))
    fl = n * flops_per_crop(model)
    res["tflops"] = {k: round(fl / min(v) / 1e9) for k, v in res["ms"].items()}
    res["speedup"] = round(min(res["ms"]["torch_eager_bf16"]) / min(res["ms"]["libvpf"]), 3)
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/torch_comparator.json", "w"), indent=1)


if __name__ == "__main__":
    main()
