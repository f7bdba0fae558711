// Fused multi-head self-attention for ViT/16 @224: T = 197 tokens, head dim 64.
// One CTA per (crop, head); Q/K/V head slices are staged once in shared memory (cp.async,
// XOR-swizzled 128-byte rows), S = Q K^T, the fp32 softmax and O = P V never leave the SM.
// Tensor work uses warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) -- attention is
// ~4 % of the ViT FLOPs; the GEMMs (gemm.cuh) carry the tcgen05 path.
//
// qkv: bf16 [n_crops*197][3*D], columns (3, H, 64) as timm's fused qkv Linear emits them.
// out: bf16 [n_crops*197][D], column h*64 + d (the layout the proj GEMM consumes).
#pragma once

#include "common.cuh"

namespace vpf {

static constexpr int kTokens = 197;
static constexpr int kHeadDim = 64;
static constexpr int kQRows = 208;    // 13 m16 tiles
static constexpr int kKeyRows = 224;  // 2 chunks x 112 keys
static constexpr int kKeyChunk = 112;
static constexpr int kAttnWarps = 7;
static constexpr int kAttnThreads = kAttnWarps * 32;
static constexpr int kAttnSmem = (kQRows + 2 * kKeyRows) * kHeadDim * 2;  // 83 968 B

__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  // byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][64] bf16 tile
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int D) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sK = sQ + kQRows * 128;
  const uint32_t sV = sK + kKeyRows * 128;
  const int head = blockIdx.x, crop = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = (size_t)3 * D;
  const __nv_bfloat16* base = qkv + (size_t)crop * kTokens * ld + head * kHeadDim;

  // ---- stage Q, K, V (rows >= 197 zero-filled) ----
  for (int i = tid; i < (kQRows + 2 * kKeyRows) * 8; i += kAttnThreads) {
    int r = i >> 3, c = i & 7;
    int which, row; uint32_t dst;
    if (r < kQRows) { which = 0; row = r; dst = sQ + swz(row, c); }
    else if (r < kQRows + kKeyRows) { which = 1; row = r - kQRows; dst = sK + swz(row, c); }
    else { which = 2; row = r - kQRows - kKeyRows; dst = sV + swz(row, c); }
    if (row < kTokens) {
      cp_async16(dst, base + (size_t)row * ld + which * D + c * 8);
    } else {
      asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int qr = lane >> 2, qc = (lane & 3) * 2;

  for (int tile = warp; tile < kQRows / 16; tile += kAttnWarps) {
    // Q fragments for the 4 k16 steps over d = 64
    uint32_t qf[4][4];
    {
      const int row = tile * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(sQ + swz(row, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      const int kc0 = ch * kKeyChunk;
      float s[14][4];
#pragma unroll
      for (int nt = 0; nt < 14; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const int krow = kc0 + nt * 8 + (lane & 7);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sK + swz(krow, kk * 4 + (lane >> 3)), b0, b1, b2, b3);
          mma_bf16_16816(s[nt], qf[kk * 2], b0, b1);
          mma_bf16_16816(s[nt], qf[kk * 2 + 1], b2, b3);
        }
      }
      // scale into the log2 domain, mask padded keys, chunk row maxima
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 14; ++nt) {
        const int col = kc0 + nt * 8 + qc;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = s[nt][e] * scale_log2;
          if (col + (e & 1) >= kTokens) v = -INFINITY;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float alpha[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);   // finite: every chunk holds a real key
        alpha[h] = exp2f(m_run[h] - m_new);
        m_run[h] = m_new;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 14; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float p = exp2f(s[nt][e] - m_run[e >> 1]);
          s[nt][e] = p;
          rs[e >> 1] += p;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * alpha[h] + rs[h];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
      }
      // O += P V over the chunk's 7 k16 steps
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        uint32_t a[4];
        {
          __nv_bfloat162 t0 = __floats2bfloat162_rn(s[2 * ks][0], s[2 * ks][1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(s[2 * ks][2], s[2 * ks][3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(s[2 * ks + 1][0], s[2 * ks + 1][1]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(s[2 * ks + 1][2], s[2 * ks + 1][3]);
          a[0] = *reinterpret_cast<uint32_t*>(&t0); a[1] = *reinterpret_cast<uint32_t*>(&t1);
          a[2] = *reinterpret_cast<uint32_t*>(&t2); a[3] = *reinterpret_cast<uint32_t*>(&t3);
        }
        const int vrow = kc0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sV + swz(vrow, dn * 2 + (lane >> 4)), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dn], a, b0, b1);
          mma_bf16_16816(o[2 * dn + 1], a, b2, b3);
        }
      }
    }
    // ---- finalise: divide by the row sums, stage through this warp's own Q rows, store ----
    float inv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float l = l_run[h];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[h] = 1.0f / l;
    }
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o[dt][0] * inv[0], o[dt][1] * inv[0]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(o[dt][2] * inv[1], o[dt][3] * inv[1]);
      const int r0 = tile * 16 + qr, r1 = r0 + 8;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r0, dt) + qc * 2),
                   "r"(*reinterpret_cast<uint32_t*>(&lo)) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r1, dt) + qc * 2),
                   "r"(*reinterpret_cast<uint32_t*>(&hi)) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;          // 16 rows x 8 chunks
      const int r = tile * 16 + (idx >> 3), c = idx & 7;
      if (r < kTokens) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + swz(r, c)));
        *reinterpret_cast<uint4*>(out + ((size_t)crop * kTokens + r) * D + head * kHeadDim + c * 8) = v;
      }
    }
  }
}

}  // namespace vpf
