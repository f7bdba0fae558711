// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M][N] = A[M][K] * W[N][K]^T  (+ fused
// epilogue), fp32 accumulation in tensor memory.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A (128 x 64) and W (BLOCK_N x 64),
//               128-byte swizzle, into a kStages-deep shared-memory ring (full/empty mbarriers)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x 4
//               per stage; tcgen05.commit releases the stage and, after the last k-block,
//               publishes the accumulator; owns the TMEM allocation (2 accumulator stages)
//   warps 2-9   epilogue: tcgen05.ld the 128 x BLOCK_N fp32 accumulator (each warp: its 32-lane
//               quarter x half of the columns), apply bias / GELU / residual / pos-embed, store.
//               Runs concurrently with the MMAs of the next tile (double-buffered accumulators).
//
// Both operands are K-major (activations [M][K] row-major; nn.Linear weights [out][in]), so no
// transposes exist anywhere.  Per output row the k-order is fixed, so results do not depend on
// how many rows (particles) a launch carries: sharding particles over GPUs is bit-invariant.
#pragma once

#include "common.cuh"
#include "sm100.cuh"

namespace vpf {

enum GemmMode : int {
  kBiasBf16 = 0,      // out bf16 [M][N] = acc + bias
  kBiasGeluBf16 = 1,  // out bf16 [M][N] = gelu(acc + bias)           (exact-erf GELU)
  kBiasResidualF32 = 2,  // X f32 [M][N] += acc + bias                (in place)
  kPatchEmbedF32 = 3  // X f32 [(r/196)*197 + 1 + r%196][N] = acc + bias + pos[1 + r%196][N]
};

struct GemmArgs {
  int M, N, K;
  const float* bias;
  void* out;
  const float* pos;
};

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
static constexpr int kUmmaK = 16;
static constexpr int kGemmThreads = 320;  // 10 warps
static constexpr int kEpiWarps = 8;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 192 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BLOCK_N > 256) ? 512 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below bf16 output rounding)
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), hx, hx);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const GemmArgs& g, int row,
                                               int col) {
  // v: 32 consecutive fp32 accumulator columns [col, col+32) of output row `row`
  const float4* b4 = reinterpret_cast<const float4*>(g.bias + col);
  if (MODE == kBiasBf16 || MODE == kBiasGeluBf16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)row * g.N + col;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 b = __ldg(b4 + q * 2 + h);
        f[h * 4 + 0] = __uint_as_float(v[q * 8 + h * 4 + 0]) + b.x;
        f[h * 4 + 1] = __uint_as_float(v[q * 8 + h * 4 + 1]) + b.y;
        f[h * 4 + 2] = __uint_as_float(v[q * 8 + h * 4 + 2]) + b.z;
        f[h * 4 + 3] = __uint_as_float(v[q * 8 + h * 4 + 3]) + b.w;
      }
      if (MODE == kBiasGeluBf16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = gelu_erf(f[e]);
      }
      uint4 o;
      o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
      o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst + q * 8) = o;
    }
  } else {
    size_t orow = (size_t)row;
    const float4* p4 = nullptr;
    if (MODE == kPatchEmbedF32) {
      int crop = row / 196, p = row - crop * 196;
      orow = (size_t)crop * 197 + 1 + p;
      p4 = reinterpret_cast<const float4*>(g.pos + (size_t)(1 + p) * g.N + col);
    }
    float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + orow * g.N + col);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 b = __ldg(b4 + q);
      float4 r = (MODE == kPatchEmbedF32) ? __ldg(p4 + q) : x4[q];
      float4 o;
      o.x = __uint_as_float(v[q * 4 + 0]) + b.x + r.x;
      o.y = __uint_as_float(v[q * 4 + 1]) + b.y + r.y;
      o.z = __uint_as_float(v[q * 4 + 2]) + b.z + r.z;
      o.w = __uint_as_float(v[q * 4 + 3]) + b.w + r.w;
      x4[q] = o;
    }
  }
}

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs g) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  uint8_t* tiles = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (g.M + kBlockM - 1) / kBlockM;
  const int n_blocks = g.N / BLOCK_N;
  const int k_blocks = g.K / kBlockK;
  const int n_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 0);
          uint8_t* sa = tiles + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1, 1);   // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 2);      // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc_k_sw128(sa);
          const uint64_t bdesc = make_smem_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (>>4) field
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                      (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);             // frees the smem stage when MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);                  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which half of the BLOCK_N columns
    constexpr int kHalfCols = BLOCK_N / 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase, 3);
      tcgen05_fence_after();
      const int row = m_blk * kBlockM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                             (uint32_t)(as * BLOCK_N + half * kHalfCols);
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (row < g.M)
          epilogue_chunk<MODE>(v, g, row, n_blk * BLOCK_N + half * kHalfCols + c * 32);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vpf
