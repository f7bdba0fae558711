// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M][N] = A[M][K] * W[N][K]^T  (+ fused
// epilogue), fp32 accumulation in tensor memory.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A (128 x 64) and W (BLOCK_N x 64),
//               128-byte swizzle, into a kStages-deep shared-memory ring (full/empty mbarriers)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x 4
//               per stage; tcgen05.commit releases the stage and, after the last k-block,
//               publishes the accumulator; owns the TMEM allocation (2 accumulator stages)
//   warps 2-9   epilogue: tcgen05.ld the 128 x BLOCK_N fp32 accumulator (each warp: its 32-lane
//               quarter x half of the columns), apply bias / GELU / residual / pos-embed, then
//               transpose through a swizzled shared-memory slab so global access is coalesced.
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
  kPatchEmbedF32 = 3,  // X f32 [(r/P)*(P+1) + 1 + r%P][N] = acc + bias + pos[1 + r%P][N], P = ppc
  // internal variant of mode 2, chosen by the launcher for small K (HBM-bound GEMMs such as proj):
  // the fp32 residual tile is loaded, updated in place and stored entirely by the TMA unit
  kBiasResidualF32Tma = 4,
  // LayerNorm folded into the GEMM that consumes it (fp16 residual stream only).  A = the residual
  // stream X itself (fp16); W' = (gamma (.) W) with each row's mean over k subtracted (stored in
  // fp16), so that x W'^T = (x - mu) (gamma (.) W)^T without knowing mu; b' = b + W beta:
  //   LN(x) W^T + b  =  r (x W'^T) + b'              (r = the row's 1/sqrt(var + eps))
  // The normalised activation is never written to or read from HBM, and the epilogue costs what the
  // plain bias epilogue costs (one FMA instead of one add).  r comes either final from
  // rowstats16_kernel or from per-column-group partial sums (sum x, sum x^2) that the producing
  // residual GEMM's epilogue leaves behind (GemmArgs::stats_out).
  kLnBiasBf16 = 5,      // out 16-bit [M][N] = r acc + b'
  kLnBiasGeluBf16 = 6   // out 16-bit [M][N] = gelu(r acc + b')
};
__host__ __device__ constexpr bool mode_out16(int m) { return m <= 1 || m == 5 || m == 6; }
__host__ __device__ constexpr bool mode_ln(int m) { return m == 5 || m == 6; }
__host__ __device__ constexpr bool mode_gelu(int m) { return m == 1 || m == 6; }

struct GemmArgs {
  int M, N, K;
  const float* bias;
  void* out;
  const float* pos;
  int f16 = 0;   // 16-bit operand / output type: 0 = bf16, 1 = fp16
  int ppc = 196; // patch-embed mode: patches per crop (196 at crop size 224); tokens = ppc + 1
  // row / ppc without a division (the epilogue needs it 32 times per thread and tile):
  // floor(n / ppc) = umulhi(n, ppc_mul) >> ppc_shift for every n < 2^31 (set_ppc); ppc == 1: n
  unsigned ppc_mul = 0, ppc_shift = 0;
  void set_ppc(int p) {
    ppc = p;
    if (p <= 1) { ppc_mul = 0; ppc_shift = 0; return; }
    int k = 0;
    while ((1 << k) < p) ++k;                              // ceil(log2 p) >= 1
    const unsigned long long two_s = 1ull << (31 + k);
    ppc_mul = (unsigned)((two_s + (unsigned long long)p - 1) / (unsigned long long)p);   // < 2^32
    ppc_shift = (unsigned)(k - 1);                         // (31 + k) - 32
  }
  int x16 = 0;   // residual modes: the residual stream X is fp16 (saturating) instead of fp32
  // LayerNorm-folded modes (5, 6): the row statistics of A.
  //   stats_parts == 0: stats[r * stats_stride] = (mu, r) final;
  //   stats_parts  > 0: stats[p * stats_ld + r * stats_stride] = (sum x, sum x^2) of column group p
  const float2* stats = nullptr;
  int stats_parts = 0;
  int stats_stride = 1;
  size_t stats_ld = 0;
  // residual modes (2, 4) with an fp16 stream: if set, the epilogue leaves the partial row sums of
  // the UPDATED (rounded) stream in stats_out[(n_blk * 2 + column half) * stats_out_ld + row]
  float2* stats_out = nullptr;
  size_t stats_out_ld = 0;
};

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
static constexpr int kUmmaK = 16;
static constexpr int kMaxEpiWarps = 16;

// Epilogue warps: 4 TMEM lane quarters x kGroups column groups.  The bf16-output epilogues
// (convert / GELU: ALU- and MUFU-heavy) get 4 warps per scheduler to hide latency; the fp32
// read-modify-write epilogues are memory-bound and need registers for loads in flight.
template <int BLOCK_N, int MODE>
struct EpiCfg {
  static constexpr int kGroups = mode_out16(MODE) ? (BLOCK_N == 192 ? 3 : 4) : 2;
  static constexpr int kWarps = 4 * kGroups;
  static constexpr int kThreads = 64 + 32 * kWarps;
  static constexpr int kColsPerWarp = BLOCK_N / kGroups;
  static_assert(kColsPerWarp % 32 == 0, "epilogue column split must be a multiple of 32");
};

template <int BLOCK_N, int CTAS, int MODE>
struct GemmCfg {
  // CTAS = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BLOCK_N tile; each CTA
  // stages its own 128 rows of A and HALF of the W tile, so shared-memory fill traffic from L2
  // per SM drops by a third versus two independent 128 x BLOCK_N tiles.
  static constexpr int kBRows = BLOCK_N / CTAS;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlabBytes = 4096;  // per epilogue warp: 32 rows x 128 B transpose slab
  // bf16 modes: one 2 KB slab per warp; residual mode: two 4 KB TMA buffers per warp (the fp32
  // residual tile is loaded, updated in place and stored by the TMA unit); patch embed: 4 KB slab
  static constexpr int kEpiBytes =
      EpiCfg<BLOCK_N, MODE>::kWarps * (mode_out16(MODE) ? 2048 : (MODE == 4 ? 8192 : 4096));
  static constexpr int kStagesMax = (227 * 1024 - kEpiBytes - 1024 - 512) / kStageBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kTmemCols = (2 * BLOCK_N > 256) ? 512 : 256;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): two fp32 lanes per instruction.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

// Exact (erf) GELU, two elements at a time, 7 instructions per element, one MUFU:
//   gelu(x) = x Phi(x) = relu(x) - |x| * 0.5 erfc(|x|/sqrt 2) = relu(x) - |x| * 2^(-q(|x|)),
//   q(a) = 1 - log2 erfc(a / sqrt 2), a degree-7 minimax fit on [0, 9] (monotone beyond; the
//   term vanishes there).  Measured against fp64 erf-GELU on [-12, 12]: abs err <= 9.6e-6,
//   rel err <= 6.9e-5 wherever |gelu| > 1e-3 -- 30x below the bf16 rounding of the output.
// Evaluated in n = -|x| (one LOP3) so that -q(|x|) = sum_k c_k (-1)^(k+1) n^k needs no negation.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 n = make_float2(-fabsf(x.x), -fabsf(x.y));
  auto k2 = [](float c) { return make_float2(c, c); };
  float2 t = ffma2(k2(9.77751013e-07f), n, k2(4.03972049e-05f));
  t = ffma2(t, n, k2(7.23606562e-04f));
  t = ffma2(t, n, k2(7.50288210e-03f));
  t = ffma2(t, n, k2(5.14463299e-02f));
  t = ffma2(t, n, k2(-4.61469725e-01f));
  t = ffma2(t, n, k2(1.14989288e+00f));
  t = ffma2(t, n, k2(-1.00009844e+00f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  return ffma2(n, e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// The same function in 8 instructions per pair, two of them MUFU (the polynomial above: 14):
//   gelu(x) = 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))),  c fitted (minimax, [-9, 9]) to the EXACT
//   erf-GELU, not the textbook tanh approximation: formula error <= 2.5e-5 absolute (textbook
//   coefficients: 4.7e-4).  tanh.approx.f32 adds <= 2^-11 relative on tanh, i.e. <= 2.4e-4 |x| on
//   the result; against fp64 erf-GELU for x ~ N(0, 1) the rms error is 1.2e-4, six times below
//   the bf16 rounding of the output (7.4e-4 rms) -- noise power +2.6 %.
// Why it matters: fc1 is the largest kernel of a frame and the only GEMM whose epilogue does real
// arithmetic; under the power cap every instruction of it is paid in clock (experiments/README.md).
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  auto k2 = [](float c) { return make_float2(c, c); };
  const float2 x2 = ffma2(x, x, k2(0.0f));
  float2 p = ffma2(k2(-3.51516788e-04f), x2, k2(3.70056460e-02f));
  p = ffma2(p, x2, k2(7.97507884e-01f));
  const float2 u = ffma2(x, p, k2(0.0f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = ffma2(x, k2(0.5f), k2(0.0f));
  return ffma2(hx, t, hx);
}
#ifndef VPF_GELU_ERF_POLY
#define VPF_GELU_ERF_POLY 0      // 1: the degree-7 erfc polynomial (A/B builds)
#endif
__device__ __forceinline__ float2 gelu2(float2 x) {
  return VPF_GELU_ERF_POLY ? gelu_erf2(x) : gelu_tanh2(x);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Epilogue of one 32-column chunk of a warp's 32-row slab.
//   phase 1 (thread = accumulator row, the tcgen05.ld 32x32b layout): bias (+GELU), convert,
//           write the row into the warp's shared-memory slab, XOR-swizzled by 16-byte piece;
//   phase 2 (transposed: consecutive lanes = consecutive 16-byte pieces of one row): read the
//           slab back and touch global memory fully coalesced (a warp instruction covers whole
//           64/128-byte row segments instead of 32 different cache lines).
// Row statistics of the LayerNorm-folded modes -> r = 1/sqrt(var + eps) of the row.
__device__ __forceinline__ float load_ln_rstd(const GemmArgs& g, int row) {
  const size_t idx = (size_t)row * (size_t)g.stats_stride;
  if (g.stats_parts == 0) {
    return __ldg(g.stats + idx).y;
  } else {
    float s = 0.f, q = 0.f;                            // summed in a fixed order: batch-invariant
    if (g.stats_parts <= 8) {                          // all loads in flight at once
      float2 v[8];
#pragma unroll
      for (int p = 0; p < 8; ++p)
        v[p] = p < g.stats_parts ? __ldg(g.stats + (size_t)p * g.stats_ld + idx) : make_float2(0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 8; ++p) { s += v[p].x; q += v[p].y; }
    } else {
      for (int p = 0; p < g.stats_parts; ++p) {
        const float2 v = __ldg(g.stats + (size_t)p * g.stats_ld + idx);
        s += v.x; q += v.y;
      }
    }
    const float inv_d = 1.0f / (float)g.K;
    const float mu = s * inv_d;
    return rsqrtf(fmaxf(fmaf(-mu, mu, q * inv_d), 0.f) + 1e-6f);
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const GemmArgs& g,
                                               const CUtensorMap* tmC, uint8_t* slab, int row0,
                                               int col, int lane, float rstd, float2 (&st)[8]) {
  const float4* b4 = reinterpret_cast<const float4*>(g.bias + col);
  if (mode_out16(MODE)) {
    // slab rows are 64 B (32 bf16): piece j of row r lives at r*64 + ((j ^ ((r>>1)&3)) << 4)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 b = __ldg(b4 + q * 2 + h);
        const float2 a01 = make_float2(__uint_as_float(v[q * 8 + h * 4 + 0]),
                                       __uint_as_float(v[q * 8 + h * 4 + 1]));
        const float2 a23 = make_float2(__uint_as_float(v[q * 8 + h * 4 + 2]),
                                       __uint_as_float(v[q * 8 + h * 4 + 3]));
        if (mode_ln(MODE)) {
          const float2 r2 = make_float2(rstd, rstd);
          f[h * 2 + 0] = ffma2(r2, a01, make_float2(b.x, b.y));
          f[h * 2 + 1] = ffma2(r2, a23, make_float2(b.z, b.w));
        } else {
          f[h * 2 + 0] = fadd2(a01, make_float2(b.x, b.y));
          f[h * 2 + 1] = fadd2(a23, make_float2(b.z, b.w));
        }
      }
      if (mode_gelu(MODE)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = gelu2(f[e]);
      }
      uint4 o;
      const bool f16 = g.f16 != 0;
      o.x = pack16x2(f[0].x, f[0].y, f16); o.y = pack16x2(f[1].x, f[1].y, f16);
      o.z = pack16x2(f[2].x, f[2].y, f16); o.w = pack16x2(f[3].x, f[3].y, f16);
      if (q == 0) {
        // the previous chunk's TMA store must have finished reading the slab (it was issued a
        // whole chunk of tcgen05.ld + math ago, so this does not stall in practice)
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      *reinterpret_cast<uint4*>(slab + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = o;
    }
    // slab (32 rows x 64 B, SWIZZLE_64B) -> global through the TMA unit: one instruction, fully
    // coalesced, rows beyond M clipped by the tensor map
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, slab, col, row0);
      tma_store_commit();
    }
    return;
  } else {
    // slab rows are 128 B (32 fp32): piece j of row r lives at r*128 + ((j ^ (r&7)) << 4)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 b = __ldg(b4 + q);
      float4 o;
      o.x = __uint_as_float(v[q * 4 + 0]) + b.x;
      o.y = __uint_as_float(v[q * 4 + 1]) + b.y;
      o.z = __uint_as_float(v[q * 4 + 2]) + b.z;
      o.w = __uint_as_float(v[q * 4 + 3]) + b.w;
      *reinterpret_cast<float4*>(slab + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
    }
    __syncwarp();
    // X element (row, col + 4 p .. +3): fp32 -> 16 bytes, fp16 residual stream -> 8 bytes
    const bool x16 = g.x16 != 0;
    uint8_t* X = reinterpret_cast<uint8_t*>(g.out);
    const size_t esz = x16 ? 2 : 4;
    const int p = lane & 7;
    float4 res[8];
    uint8_t* dst[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {           // all global loads first: 8 independent requests
      const int r = i * 4 + (lane >> 3);
      const int grow = row0 + r;
      dst[i] = nullptr;
      if (grow < g.M) {
        if (MODE == kPatchEmbedF32) {
          const int crop = g.ppc_mul ? (int)(__umulhi((unsigned)grow, g.ppc_mul) >> g.ppc_shift) : grow;
          const int pp = grow - crop * g.ppc;
          dst[i] = X + ((((size_t)crop * (g.ppc + 1) + 1 + pp) * g.N + col) + 4 * p) * esz;
          res[i] = __ldg(reinterpret_cast<const float4*>(g.pos + (size_t)(1 + pp) * g.N + col) + p);
        } else {
          dst[i] = X + (((size_t)grow * g.N + col) + 4 * p) * esz;
          if (x16) {
            const uint2 h = *reinterpret_cast<const uint2*>(dst[i]);
            const float2 lo = unpack_h2(h.x), hi = unpack_h2(h.y);
            res[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            res[i] = *reinterpret_cast<const float4*>(dst[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const float4 a = *reinterpret_cast<const float4*>(slab + r * 128 + ((p ^ (r & 7)) << 4));
      if (dst[i]) {
        float4 o;
        o.x = a.x + res[i].x; o.y = a.y + res[i].y; o.z = a.z + res[i].z; o.w = a.w + res[i].w;
        if (x16) {
          const uint2 pk = make_uint2(pack_h2_sat(o.x, o.y), pack_h2_sat(o.z, o.w));
          *reinterpret_cast<uint2*>(dst[i]) = pk;
          if (MODE == kBiasResidualF32 && g.stats_out) {   // sums of the stream as stored
            const float2 y0 = unpack_h2(pk.x), y1 = unpack_h2(pk.y);
            st[i].x += (y0.x + y0.y) + (y1.x + y1.y);
            st[i].y = fmaf(y0.x, y0.x, fmaf(y0.y, y0.y, fmaf(y1.x, y1.x, fmaf(y1.y, y1.y, st[i].y))));
          }
        } else {
          *reinterpret_cast<float4*>(dst[i]) = o;
        }
      }
    }
  }
  __syncwarp();   // slab is rewritten by the next chunk
}

template <int BLOCK_N, int MODE, int CTAS>
__global__ void __launch_bounds__((EpiCfg<BLOCK_N, MODE>::kThreads), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using Cfg = GemmCfg<BLOCK_N, CTAS, MODE>;
  using Epi = EpiCfg<BLOCK_N, MODE>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles (identical offsets in both CTAs of a pair)
  uint8_t* tiles = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* slabs = tiles + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* epi_bar = full_bar + 32;   // [epilogue warp][2] "residual tile landed" (mode 2)

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int m_blocks = (g.M + kBlockM * CTAS - 1) / (kBlockM * CTAS);   // per CTA group
  const int n_blocks = g.N / BLOCK_N;
  const int k_blocks = g.K / kBlockK;
  const int n_tiles = m_blocks * n_blocks;
  const int group = blockIdx.x / CTAS, n_groups = gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], Epi::kWarps * CTAS);   // epilogue warps of every CTA of the group
    }
    if (MODE == kBiasResidualF32Tma)
      for (int s = 0; s < 2 * Epi::kWarps; ++s) mbar_init(&epi_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();      // everything above overlapped the previous kernel's tail; operands are its output

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows / W half) ==========
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = group; tile < n_tiles; tile += n_groups) {
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        const int a_row = (m_blk * CTAS + (int)cta_rank) * kBlockM;
        const int b_row = n_blk * BLOCK_N + (int)cta_rank * Cfg::kBRows;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait<true>(&empty_bar[stage], phase ^ 1, 0);
          uint8_t* sa = tiles + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (CTAS == 2) {
            // bytes of both CTAs are counted on the leader's barrier (its MMA thread waits there)
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * 2);
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(sa, &tmA, bar, kb * kBlockK, a_row);
            tma_load_2d_pair(sb, &tmB, bar, kb * kBlockK, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBlockK, a_row);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, b_row);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      // folded-LayerNorm modes: A is the fp16 residual stream, so W' is stored in fp16 whatever the
      // encoder's operand type (kind::f16 with A = f16, B = bf16 raises an illegal-instruction
      // fault on sm_100a: measured, round 2); g.f16 then only selects the OUTPUT's 16-bit type
      const uint32_t idesc = make_idesc_16(kBlockM * CTAS, BLOCK_N, mode_ln(MODE) || g.f16 != 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait<true>(&tmem_empty[as], aphase ^ 1, 1);   // epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait<true>(&full_bar[stage], phase, 2);      // TMA bytes (of both CTAs) have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc_k_sw128(sa);
          const uint64_t bdesc = make_smem_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (>>4) field
            if (CTAS == 2)
              umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                             (uint32_t)((kb | k) != 0));
            else
              umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                        (uint32_t)((kb | k) != 0));
          }
          // frees the smem stage (in both CTAs) when the MMAs retire
          if (CTAS == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogues (of both CTAs)
        if (CTAS == 2) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ===================== epilogue warps 2.. =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which column group of the BLOCK_N columns
    constexpr int kHalfCols = Epi::kColsPerWarp;
    constexpr int kSlab = mode_out16(MODE) ? 2048 : 4096;
    int it = 0;
    if constexpr (MODE == kBiasResidualF32Tma) {
      // ---- fp32 residual, in place, entirely through the TMA unit -------------------------
      // Each warp owns two 4 KB buffers (32 rows x 32 fp32, 128-byte swizzle).  The residual
      // tile of chunk j+1 is in flight while chunk j is updated in place (x += acc + bias, thread
      // = row, conflict-free) and handed back to the TMA unit for the store.  No global address
      // arithmetic, no register staging, rows beyond M clipped / zero-filled by the tensor map.
      constexpr int kChunks = kHalfCols / 32;
      uint8_t* xbuf = slabs + (warp - 2) * 8192;
      uint64_t* xbar = epi_bar + (warp - 2) * 2;
      uint32_t j = 0;
      const bool x16 = g.x16 != 0;           // fp16 residual: 32 x 32 tile = 2 KB, 64-byte swizzle
      auto issue_load = [&](int mb, int nb, int c_, uint32_t jj) {   // (row block, column block)
        mbar_arrive_expect_tx(&xbar[jj & 1], x16 ? 2048u : 4096u);
        tma_load_2d(xbuf + (jj & 1) * 4096, &tmC, &xbar[jj & 1],
                    nb * BLOCK_N + half * kHalfCols + c_ * 32,
                    (mb * CTAS + (int)cta_rank) * kBlockM + quarter * 32);
      };
      if (lane == 0 && group < n_tiles) issue_load(group / n_blocks, group % n_blocks, 0, 0);
      for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait<true>(&tmem_full[as], aphase, 3);
        tcgen05_fence_after();
        const int row0 = (m_blk * CTAS + (int)cta_rank) * kBlockM + quarter * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                               (uint32_t)(as * BLOCK_N + half * kHalfCols);
        float2 ssum = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f);   // this row, this column group
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c, ++j) {
          if (lane == 0) {                       // prefetch the next chunk's residual tile
            // next chunk of this tile, or the first chunk of the next tile (one integer division
            // per tile, not per chunk)
            int nm = m_blk, nn = n_blk, nc = c + 1;
            bool more = true;
            if (nc == kChunks) {
              const int nt = tile + n_groups;
              more = nt < n_tiles;
              nm = nt / n_blocks; nn = nt - nm * n_blocks; nc = 0;
            }
            if (more) {
              tma_store_wait_read<0>();          // the other buffer's last store has left smem
              issue_load(nm, nn, nc, j + 1);
            }
          }
          __syncwarp();
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          const int col = n_blk * BLOCK_N + half * kHalfCols + c * 32;
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + col);
          mbar_wait<true>(&xbar[j & 1], (j >> 1) & 1, 4);
          if (x16) {
            // row = 64 B: 16-byte piece q (8 halves) at lane*64 + ((q ^ ((lane>>1)&3)) << 4)
            uint8_t* buf = xbuf + (j & 1) * 4096 + lane * 64;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4* px = reinterpret_cast<uint4*>(buf + ((q ^ ((lane >> 1) & 3)) << 4));
              const float4 b0 = __ldg(b4 + 2 * q), b1 = __ldg(b4 + 2 * q + 1);
              uint4 h = *px;
              float2 x0 = unpack_h2(h.x), x1 = unpack_h2(h.y), x2 = unpack_h2(h.z), x3 = unpack_h2(h.w);
              h.x = pack_h2_sat(x0.x + (__uint_as_float(v[q * 8 + 0]) + b0.x),
                                x0.y + (__uint_as_float(v[q * 8 + 1]) + b0.y));
              h.y = pack_h2_sat(x1.x + (__uint_as_float(v[q * 8 + 2]) + b0.z),
                                x1.y + (__uint_as_float(v[q * 8 + 3]) + b0.w));
              h.z = pack_h2_sat(x2.x + (__uint_as_float(v[q * 8 + 4]) + b1.x),
                                x2.y + (__uint_as_float(v[q * 8 + 5]) + b1.y));
              h.w = pack_h2_sat(x3.x + (__uint_as_float(v[q * 8 + 6]) + b1.z),
                                x3.y + (__uint_as_float(v[q * 8 + 7]) + b1.w));
              *px = h;
              if (g.stats_out) {                       // sums of the stream as stored (rounded)
                x0 = unpack_h2(h.x); x1 = unpack_h2(h.y); x2 = unpack_h2(h.z); x3 = unpack_h2(h.w);
                ssum = fadd2(ssum, fadd2(fadd2(x0, x1), fadd2(x2, x3)));
                ssq = ffma2(x0, x0, ssq); ssq = ffma2(x1, x1, ssq);
                ssq = ffma2(x2, x2, ssq); ssq = ffma2(x3, x3, ssq);
              }
            }
          } else {
            uint8_t* buf = xbuf + (j & 1) * 4096 + lane * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4* px = reinterpret_cast<float4*>(buf + ((q ^ (lane & 7)) << 4));
              const float4 b = __ldg(b4 + q);
              float4 x = *px;
              x.x += __uint_as_float(v[q * 4 + 0]) + b.x;
              x.y += __uint_as_float(v[q * 4 + 1]) + b.y;
              x.z += __uint_as_float(v[q * 4 + 2]) + b.z;
              x.w += __uint_as_float(v[q * 4 + 3]) + b.w;
              *px = x;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, xbuf + (j & 1) * 4096, col, row0);
            tma_store_commit();
          }
        }
        if (g.stats_out && row0 + lane < g.M)
          g.stats_out[(size_t)(n_blk * Epi::kGroups + half) * g.stats_out_ld + (size_t)(row0 + lane)] =
              make_float2(ssum.x + ssum.y, ssq.x + ssq.y);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
          else mbar_arrive(&tmem_empty[as]);
        }
      }
      if (lane == 0) tma_store_wait_read<0>();
    } else {
    for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = (m_blk * CTAS + (int)cta_rank) * kBlockM + quarter * 32;
      float rstd = 1.f;
      if constexpr (mode_ln(MODE)) {     // thread = accumulator row; rows past M are clipped by TMA
        // fetched while this tile's MMAs are still running: the latency hides behind the wait
        const int r = row0 + lane;
        rstd = load_ln_rstd(g, r < g.M ? r : g.M - 1);
      }
      mbar_wait<true>(&tmem_full[as], aphase, 3);
      tcgen05_fence_after();
      uint8_t* slab = slabs + (warp - 2) * kSlab;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                             (uint32_t)(as * BLOCK_N + half * kHalfCols);
      float2 st[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) st[i] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        epilogue_chunk<MODE>(v, g, &tmC, slab, row0, n_blk * BLOCK_N + half * kHalfCols + c * 32, lane,
                             rstd, st);
      }
      if (MODE == kBiasResidualF32 && g.stats_out) {
        // transposed phase: lanes 8 r' .. 8 r' + 7 hold the 4-column pieces of row i*4 + r'
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float s = st[i].x, q = st[i].y;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
          }
          const int r = row0 + i * 4 + (lane >> 3);
          if ((lane & 7) == 0 && r < g.M)
            g.stats_out[(size_t)(n_blk * Epi::kGroups + half) * g.stats_out_ld + (size_t)r] =
                make_float2(s, q);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
        else mbar_arrive(&tmem_empty[as]);
      }
    }
    }
    if (mode_out16(MODE) && lane == 0) tma_store_wait_read<0>();   // smem must outlive the last store
  }

  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vpf
