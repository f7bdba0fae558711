"""Per-kernel timings of the ViT stages (CUDA events, L2 flushed by rotating large buffers)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    dev = torch.device("cuda")
    M = n * 197
    H = D // 64
    res = {"n_crops": n, "D": D, "M": M}
    x = torch.randn(M, D, device=dev)
    xn = torch.randn(M, D, device=dev).to(torch.bfloat16)
    big = torch.randn(M, 4 * D, device=dev).to(torch.bfloat16)
    g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    shapes = {"qkv": (3 * D, D, 0, xn), "proj": (D, D, 2, xn), "fc1": (4 * D, D, 1, xn),
              "fc2": (D, 4 * D, 2, big)}
    for name, (N, K, mode, a) in shapes.items():
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.zeros(N, device=dev)
        out = x if mode == 2 else torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.gemm_bf16(a, w, bias, mode, out=out))
        res[name] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
        ref = timeit(lambda: torch.matmul(a, w.t()))
        res[name]["cublas_ms"] = ref
        res[name]["cublas_tflops"] = 2.0 * M * N * K / ref / 1e9
        del w, out
    qkv = big[:, :3 * D].contiguous()
    ms = timeit(lambda: ops.attention(qkv, n, H))
    res["attention"] = {"ms": ms, "tflops": 4.0 * 197 * 197 * D * n / ms / 1e9}
    ms = timeit(lambda: ops.layernorm(x, g, b))
    res["layernorm"] = {"ms": ms, "gbs": M * D * 6 / ms / 1e6}
    frame = torch.randint(0, 256, (1080, 1920, 3), dtype=torch.uint8, device=dev)
    st = torch.zeros(6, n, device=dev); st[0] = 960; st[1] = 540
    st[0] += torch.randn(n, device=dev) * 20
    patches = torch.empty(n * 196, 768, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.crop_state(frame, st, 96.0, 96.0, patches))
    res["crop"] = {"ms": ms, "gbs": n * 301056 / ms / 1e6}
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/kernels_n{n}_d{D}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
