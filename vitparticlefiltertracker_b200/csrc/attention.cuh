// Fused multi-head self-attention for ViT/16, head dim 64, any token count T <= 208 -- the
// warp-level (mma.sync m16n8k16, fp32 accumulate) kernels.  The default 224-pixel crops
// (T = 197) run on the tcgen05 kernel in attention_tc.cuh; these serve the smaller crop sizes
// ("native-resolution crops": T = (S/16)^2 + 1 tokens):
//
//   attention_small_kernel<KT>  T <= KT in {32, 64}: one WARP per (crop, head); Q/K/V head slices
//                               staged in the warp's own shared-memory region, one-pass softmax
//   attention_kernel            64 < T <= 208 (and VPF_ATTN=mma at T = 197): one CTA per (crop,
//                               head), 7 warps over the query tiles, keys in two chunks of 112
//                               with an online softmax
//
// S = Q K^T, the fp32 softmax and O = P V never leave the SM.
// qkv: bf16/fp16 [n_crops*T][3*D], columns (3, H, 64) as timm's fused qkv Linear emits them.
// out: bf16/fp16 [n_crops*T][D], column h*64 + d (the layout the proj GEMM consumes).
#pragma once

#include "common.cuh"

namespace vpf {

static constexpr int kMaxTokens = 197;  // 224-pixel crops; smaller crop sizes give fewer tokens
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

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                              uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// operand type is a warp-uniform run-time flag (bf16 default, fp16 optional)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1, bool f16) {
  if (f16) mma_f16_16816(d, a, b0, b1); else mma_bf16_16816(d, a, b0, b1);
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int D,
                 int T /* tokens, <= 208 */, bool f16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sK = sQ + kQRows * 128;
  const uint32_t sV = sK + kKeyRows * 128;
  pdl_launch_dependents();
  // (crop, head) flattened into grid.x: gridDim.y is capped at 65535, n_crops is not
  const int heads_ = D / kHeadDim;
  const int crop = blockIdx.x / heads_, head = blockIdx.x - crop * heads_;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = (size_t)3 * D;
  pdl_wait();
  const __nv_bfloat16* base = qkv + (size_t)crop * T * ld + head * kHeadDim;

  // ---- stage Q, K, V (rows >= T zero-filled) ----
  for (int i = tid; i < (kQRows + 2 * kKeyRows) * 8; i += kAttnThreads) {
    int r = i >> 3, c = i & 7;
    int which, row; uint32_t dst;
    if (r < kQRows) { which = 0; row = r; dst = sQ + swz(row, c); }
    else if (r < kQRows + kKeyRows) { which = 1; row = r - kQRows; dst = sK + swz(row, c); }
    else { which = 2; row = r - kQRows - kKeyRows; dst = sV + swz(row, c); }
    if (row < T) {
      cp_async16(dst, base + (size_t)row * ld + which * D + c * 8);
    } else {
      asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int qr = lane >> 2, qc = (lane & 3) * 2;

  const int q_tiles = (T + 15) >> 4;
  const int key_chunks = T > kKeyChunk ? 2 : 1;   // a chunk without a real key is skipped
  for (int tile = warp; tile < q_tiles; tile += kAttnWarps) {
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
    for (int ch = 0; ch < key_chunks; ++ch) {
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
          mma_16816(s[nt], qf[kk * 2], b0, b1, f16);
          mma_16816(s[nt], qf[kk * 2 + 1], b2, b3, f16);
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
          if (col + (e & 1) >= T) v = -INFINITY;
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
        a[0] = pack16x2(s[2 * ks][0], s[2 * ks][1], f16);
        a[1] = pack16x2(s[2 * ks][2], s[2 * ks][3], f16);
        a[2] = pack16x2(s[2 * ks + 1][0], s[2 * ks + 1][1], f16);
        a[3] = pack16x2(s[2 * ks + 1][2], s[2 * ks + 1][3], f16);
        const int vrow = kc0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sV + swz(vrow, dn * 2 + (lane >> 4)), b0, b1, b2, b3);
          mma_16816(o[2 * dn], a, b0, b1, f16);
          mma_16816(o[2 * dn + 1], a, b2, b3, f16);
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
      const uint32_t lo = pack16x2(o[dt][0] * inv[0], o[dt][1] * inv[0], f16);
      const uint32_t hi = pack16x2(o[dt][2] * inv[1], o[dt][3] * inv[1], f16);
      const int r0 = tile * 16 + qr, r1 = r0 + 8;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r0, dt) + qc * 2), "r"(lo) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r1, dt) + qc * 2), "r"(hi) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;          // 16 rows x 8 chunks
      const int r = tile * 16 + (idx >> 3), c = idx & 7;
      if (r < T) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + swz(r, c)));
        *reinterpret_cast<uint4*>(out + ((size_t)crop * T + r) * D + head * kHeadDim + c * 8) = v;
      }
    }
  }
}

// ---- T <= KT (32 or 64): one warp per (crop, head) ----------------------------------------------
// The warp stages its pair's Q, K, V slices ([KT][64] each, 128-byte rows, XOR-swizzled, rows >= T
// zero-filled) in its own shared-memory region, then for every 16-row query tile: S = Q K^T over
// all KT keys at once (no online rescaling), fp32 softmax in the log2 domain, O = P V, and a
// coalesced store staged through the tile's own Q rows.  No block-level synchronisation at all.
static constexpr int kSmallWarps = 4;
template <int KT>
struct SmallAttn {
  static constexpr int kWarpBytes = 3 * KT * 128;
  static constexpr int kSmem = kSmallWarps * kWarpBytes;
};

template <int KT>
__global__ void __launch_bounds__(kSmallWarps * 32)
attention_small_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                       int n_pairs, int heads, int T, bool f16) {
  static_assert(KT == 32 || KT == 64, "padded token count");
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kSmallWarps + warp;
  if (pair >= n_pairs) return;                        // warp-uniform; no block barriers below
  pdl_wait();
  const int crop = pair / heads, head = pair - crop * heads;
  const int D = heads * kHeadDim;
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(smem) + warp * SmallAttn<KT>::kWarpBytes;
  const uint32_t sK = sQ + KT * 128;
  const uint32_t sV = sK + KT * 128;
  const size_t ld = (size_t)3 * D;
  const __nv_bfloat16* base = qkv + (size_t)crop * T * ld + head * kHeadDim;

#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const uint32_t dst0 = sQ + which * KT * 128;
#pragma unroll
    for (int i = lane; i < KT * 8; i += 32) {
      const int r = i >> 3, c = i & 7;
      if (r < T) cp_async16(dst0 + swz(r, c), base + (size_t)r * ld + which * D + c * 8);
      else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst0 + swz(r, c)), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int qr = lane >> 2, qc = (lane & 3) * 2;
  constexpr int kNT = KT / 8;                              // key n8 tiles
  const int q_tiles = (T + 15) >> 4;
  const int key_tiles = (T + 7) >> 3;                      // n8 tiles that hold a real key

#pragma unroll 1
  for (int tile = 0; tile < q_tiles; ++tile) {
    uint32_t qf[4][4];
    {
      const int row = tile * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(sQ + swz(row, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float s[kNT][4];
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt < key_tiles) {                                // warp-uniform
        const int krow = nt * 8 + (lane & 7);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sK + swz(krow, kk * 4 + (lane >> 3)), b0, b1, b2, b3);
          mma_16816(s[nt], qf[kk * 2], b0, b1, f16);
          mma_16816(s[nt], qf[kk * 2 + 1], b2, b3, f16);
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const int col = nt * 8 + qc;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[nt][e] * scale_log2;
        if (col + (e & 1) >= T) v = -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {                          // key 0 is always real: maxima are finite
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - mx[e >> 1]);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KT / 16; ++ks) {
      if (ks * 16 < T) {                                   // warp-uniform
        uint32_t a[4];
        a[0] = pack16x2(s[2 * ks][0], s[2 * ks][1], f16);
        a[1] = pack16x2(s[2 * ks][2], s[2 * ks][3], f16);
        a[2] = pack16x2(s[2 * ks + 1][0], s[2 * ks + 1][1], f16);
        a[3] = pack16x2(s[2 * ks + 1][2], s[2 * ks + 1][3], f16);
        const int vrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sV + swz(vrow, dn * 2 + (lane >> 4)), b0, b1, b2, b3);
          mma_16816(o[2 * dn], a, b0, b1, f16);
          mma_16816(o[2 * dn + 1], a, b2, b3, f16);
        }
      }
    }
    float inv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float l = rs[h];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[h] = 1.0f / l;
    }
    __syncwarp();                                          // every lane has read its Q fragments
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const uint32_t lo = pack16x2(o[dt][0] * inv[0], o[dt][1] * inv[0], f16);
      const uint32_t hi = pack16x2(o[dt][2] * inv[1], o[dt][3] * inv[1], f16);
      const int r0 = tile * 16 + qr, r1 = r0 + 8;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r0, dt) + qc * 2), "r"(lo) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r1, dt) + qc * 2), "r"(hi) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;                       // 16 rows x 8 chunks
      const int r = tile * 16 + (idx >> 3), c = idx & 7;
      if (r < T) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + swz(r, c)));
        *reinterpret_cast<uint4*>(out + ((size_t)crop * T + r) * D + head * kHeadDim + c * 8) = v;
      }
    }
  }
}

}  // namespace vpf
