// Shared helpers for libvpf.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vpf.h"

namespace vpf {

// thread-local last-error message, read through vpf_last_error()
void set_error(const char* fmt, ...);

#define VPF_REQUIRE(cond, ...)                                                     \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      ::vpf::set_error(__VA_ARGS__);                                               \
      return VPF_ERR_INVALID;                                                      \
    }                                                                              \
  } while (0)

#define VPF_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t err__ = (call);                                                    \
    if (err__ != cudaSuccess) {                                                    \
      ::vpf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
                       cudaGetErrorString(err__));                                 \
      return VPF_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define VPF_LAUNCH_CHECK() VPF_CUDA(cudaGetLastError())

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Programmatic Dependent Launch ------------------------------------------------------------
// Every kernel of the per-frame chain calls pdl_launch_dependents() first (the next kernel of the
// stream may then be scheduled onto SMs as this grid drains, and run its prologue: barrier init,
// tensor-memory allocation, descriptor prefetch) and pdl_wait() before its first global-memory
// access (returns once the preceding grid has completed and its writes are visible), so the chain
// keeps plain stream-order semantics.  Kernels launched without the attribute behave as before.
// Measured (tools/bench_pdl.py, experiments/README.md): -7.5 % per frame for ViT-Ti / 256 particles
// (short kernels, prologue-bound) but +2-3 % for the 4096-particle ViT-B shapes, so the attribute
// is only set for small launches (pdl_for_rows).
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// rows = token (or patch) rows the launch works on.  VPF_PDL=0 / 1 forces the attribute off / on
// for A/B runs; default: on up to kPdlMaxRows rows.  Defined in vit.cu.
static constexpr int kPdlMaxRows = 65536;
bool pdl_for_rows(long long rows);

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 16-bit operand type of the encoder: bf16 (default) or fp16 (same tensor-core rate and bytes,
// 3 more mantissa bits, less range) -- a run-time, warp-uniform flag.
__device__ __forceinline__ uint32_t pack16x2(float a, float b, bool f16) {
  if (f16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t v, bool f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
__device__ __forceinline__ uint16_t pack16(float a, bool f16) {
  if (f16) { __half h = __float2half_rn(a); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<uint16_t*>(&h);
}

// ---- 16-bit (fp16) residual stream option: saturating conversions ------------------------------
// two floats -> packed fp16x2 (first element in the low half), overflow clamps to +-65504
// -DVPF_EXPERIMENT_STREAM_BF16=1 (A/B builds only, tools/ab_bench.sh): the stream keeps bf16's 8
// significant bits inside its fp16 container -- a measurement of what the mantissa width of the
// folded GEMMs' A operand costs under the power cap, NOT a shippable mode (it coarsens the
// residual stream itself).
#ifndef VPF_EXPERIMENT_STREAM_BF16
#define VPF_EXPERIMENT_STREAM_BF16 0
#endif
__device__ __forceinline__ uint32_t pack_h2_sat(float first, float second) {
#if VPF_EXPERIMENT_STREAM_BF16
  first = __bfloat162float(__float2bfloat16_rn(first));
  second = __bfloat162float(__float2bfloat16_rn(second));
#endif
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first));
  return d;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// ---- Philox4x32-10 (Random123 constants), shared by host and device ----------------------
struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

}  // namespace vpf
