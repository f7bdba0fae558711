// Fused multi-head self-attention for ViT/16 @224 on tcgen05 (T = 197 tokens, head dim 64).
//
// Persistent, one CTA per SM, looping over (crop, head) pairs.  A pair is two "units" (query
// tiles of 128 rows: tokens 0-127 and 128-196); each unit owns one TMEM buffer and one softmax
// warpgroup, so the tensor pipe works for one unit while the other unit's softmax runs:
//
//   warp 0      TMA producer: Q tile (128 x 64) per unit, K and V (208 x 64) per pair, from a
//               3-D tensor map [crop][token][3D] so tokens >= 197 are zero-filled by the TMA
//               unit itself; 128-byte swizzle; 4 Q stages / 2 KV stages.
//   warp 1      MMA issuer (one thread):  S = Q K^T   tcgen05.mma M=128 N=208 K=16 x4 (SS)
//                                         O = P V     tcgen05.mma M=128 N=64  K=16 x13 with P
//               read straight from TMEM (TS form) and V as an MN-major shared-memory operand.
//   warps 2-5   softmax warpgroup of tile 0 / TMEM buffer 0; warps 6-9 the same for tile 1:
//               thread = query row; two tcgen05.ld passes over S (row max, then exp2 + row sum),
//               P written back in place as packed bf16 with tcgen05.st; after O = P V the same
//               threads normalise O by the row sum and store it through the (now free) Q stage,
//               coalesced.
//
// TMEM (512 columns): buffer b at column 256 b:  S [0,208)  P aliases [0,104)  O [128,192).
// qkv: bf16 [n_crops*197][3D] (timm's fused qkv layout), out: bf16 [n_crops*197][D].
#pragma once

#include "common.cuh"
#include "sm100.cuh"

// 1: half of the softmax exponentials on the FMA pipe (ex2_poly2); 0 (default): all on the MUFU
// unit.  Measured (round 2, same-box A/B with in-step kernel timing, profiles/ab_attn_poly_r02.txt):
// the polynomial makes the kernel 14 % FASTER alone (boost clocks: MUFU-bound) but 21-27 % SLOWER
// inside the tracking step (16.5 -> 20-21 ms per frame): the encoder's GEMMs hold the SM clock near
// 1.2 GHz there, and at that clock the kernel is bound by instruction issue, which the polynomial
// adds to.  Kept as a build-time option only.
#ifndef VPF_ATTN_POLY_EXP2
#define VPF_ATTN_POLY_EXP2 0
#endif

namespace vpf {
namespace attn {

static constexpr int kT = 197;          // tokens
static constexpr int kD = 64;           // head dim
static constexpr int kKeys = 208;       // keys padded to a multiple of 16 (UMMA N / K granularity)
static constexpr int kQStages = 4;
static constexpr int kKVStages = 2;
static constexpr int kQBytes = 128 * kD * 2;        // 16 KB
static constexpr int kKVBytes = kKeys * kD * 2;     // 26 KB (1024-byte multiple)
static constexpr int kThreads = 320;
// the run-time-token kernel has a second MMA-issuing warp (one per TMEM buffer)
static constexpr int kThreadsVar = 352;
static constexpr int kMmaWarp1 = 10;
static constexpr int kSmemBytes = kQStages * kQBytes + kKVStages * 2 * kKVBytes + 1024 + 256;
static constexpr int kTmemCols = 512;
static constexpr int kBufCols = 256;
static constexpr int kOCol = 128;

// tcgen05.mma with the A operand in tensor memory (TS form)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major operand ([k][64 contiguous n] rows of 128 B, 128-byte swizzle, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused (a single 64-element group along N)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: next group of 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bmn(int M, int N, bool f16) {
  return make_idesc_16(M, N, f16) | (1u << 16);     // B operand MN-major
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 ffma2v(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2v(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, bool f16) { return pack16x2(a, b, f16); }

// 2^x for two values x <= 0 on the FMA pipe (no MUFU): round-to-nearest split x = n + f with the
// 1.5 * 2^23 magic constant, degree-4 minimax polynomial for 2^f on [-0.5, 0.5] (relative error
// 2.7e-6 in fp32 Horner form: 200x below the bf16 / 16x below the fp16 rounding of P), then n added
// to the exponent field.  The softmax is bound by the 16 exp2 lanes per clock of the MUFU unit
// (profiles/ncu_crop_attn_r02.json: XU pipe the busiest at 47 %, FMA / issue slots at 37 %), so half
// of the exponentials of every chunk are computed here instead.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);                       // 2^-125: still a normal number after the shift
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 r = fadd2v(x, magic);               // low mantissa bits = round(x)
  const float2 n = fadd2v(r, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = fadd2v(x, make_float2(-n.x, -n.y));
  float2 p = ffma2v(make_float2(9.570101276e-03f, 9.570101276e-03f), f,
                    make_float2(5.591785908e-02f, 5.591785908e-02f));
  p = ffma2v(p, f, make_float2(2.402474433e-01f, 2.402474433e-01f));
  p = ffma2v(p, f, make_float2(6.931217909e-01f, 6.931217909e-01f));
  p = ffma2v(p, f, make_float2(9.999992847e-01f, 9.999992847e-01f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23)));
}

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    __nv_bfloat16* __restrict__ out, int n_pairs, int heads, int D, int f16_) {
  const bool f16 = f16_ != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = base;                                   // [kQStages][128][128 B]
  uint8_t* sK = sQ + kQStages * kQBytes;                // [kKVStages][208][128 B]
  uint8_t* sV = sK + kKVStages * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kKVBytes);
  uint64_t* q_full = bars;            // [4]
  uint64_t* q_empty = bars + 4;       // [4]
  uint64_t* kv_full = bars + 8;       // [2]
  uint64_t* kv_empty = bars + 10;     // [2]
  uint64_t* s_full = bars + 12;       // [2] per TMEM buffer
  uint64_t* p_full = bars + 14;       // [2]
  uint64_t* o_full = bars + 16;       // [2]
  uint64_t* buf_empty = bars + 18;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_pairs = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < kQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 4); }
    for (int i = 0; i < kKVStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&buf_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int ip = 0; ip < my_pairs; ++ip) {
        const int pair = (int)blockIdx.x + ip * (int)gridDim.x;
        const int crop = pair / heads, head = pair - crop * heads;
        const int ks = ip & 1;
        mbar_wait(&kv_empty[ks], ((ip >> 1) & 1) ^ 1, 10);
        mbar_arrive_expect_tx(&kv_full[ks], 2 * kKVBytes);
        tma_load_3d(sK + ks * kKVBytes, &tmKV, &kv_full[ks], D + head * kD, 0, crop);
        tma_load_3d(sV + ks * kKVBytes, &tmKV, &kv_full[ks], 2 * D + head * kD, 0, crop);
        for (int t = 0; t < 2; ++t) {
          const int u = 2 * ip + t, qs = u & 3;
          mbar_wait(&q_empty[qs], ((u >> 2) & 1) ^ 1, 11);
          mbar_arrive_expect_tx(&q_full[qs], kQBytes);
          tma_load_3d(sQ + qs * kQBytes, &tmQ, &q_full[qs], head * kD, t * 128, crop);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_16(128, kKeys, f16);
      const uint32_t idesc_o = make_idesc_bmn(128, kD, f16);
      const int units = 2 * my_pairs;
      auto issue_s = [&](int u) {
        const int ip = u >> 1, b = u & 1, qs = u & 3, ks = ip & 1;
        mbar_wait(&q_full[qs], (u >> 2) & 1, 12);
        if (b == 0) mbar_wait(&kv_full[ks], (ip >> 1) & 1, 13);
        mbar_wait(&buf_empty[b], (ip & 1) ^ 1, 14);
        tcgen05_fence_after();
        const uint64_t adesc = make_smem_desc_k_sw128(smem_u32(sQ + qs * kQBytes));
        const uint64_t bdesc = make_smem_desc_k_sw128(smem_u32(sK + ks * kKVBytes));
        const uint32_t d = tmem_base + (uint32_t)(b * kBufCols);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
        umma_commit(&s_full[b]);
      };
      auto issue_pv = [&](int u) {
        const int ip = u >> 1, b = u & 1, ks = ip & 1;
        mbar_wait(&p_full[b], ip & 1, 15);
        tcgen05_fence_after();
        const uint32_t vaddr = smem_u32(sV + ks * kKVBytes);
        const uint32_t p = tmem_base + (uint32_t)(b * kBufCols);
        const uint32_t d = p + (uint32_t)kOCol;
#pragma unroll
        for (int k = 0; k < kKeys / 16; ++k)      // 16 keys: 8 packed P columns, 2 V row atoms
          umma_bf16_ts(d, p + (uint32_t)(k * 8), make_smem_desc_mn_sw128(vaddr + k * 2048), idesc_o,
                       (uint32_t)(k != 0));
        umma_commit(&o_full[b]);
        if (b == 1) umma_commit(&kv_empty[ks]);   // every MMA of this pair has retired
      };
      if (units > 0) { issue_s(0); issue_s(1); }
      for (int u = 0; u < units; ++u) {
        issue_pv(u);
        if (u + 2 < units) issue_s(u + 2);
      }
    }
  } else {
    // ===================== softmax / output warpgroups =====================
    const int b = (warp - 2) >> 2;                 // tile index == TMEM buffer
    const int quarter = warp & 3;                  // TMEM lane quarter
    const int row = quarter * 32 + lane;           // row inside the 128-row tile
    const int token = b * 128 + row;
    const bool warp_has_rows = (b * 128 + quarter * 32) < kT;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t tbuf = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * kBufCols);
    for (int ip = 0; ip < my_pairs; ++ip) {
      const int pair = (int)blockIdx.x + ip * (int)gridDim.x;
      const int crop = pair / heads, head = pair - crop * heads;
      const int u = 2 * ip + b, qs = u & 3;
      mbar_wait(&s_full[b], ip & 1, 16);
      tcgen05_fence_after();
      float inv_l = 0.f;
      if (warp_has_rows) {
        // Both passes are software-pipelined: the tcgen05.ld of chunk c+1 is in flight while
        // chunk c is processed (two register buffers).  Chunks: 6 x 32 columns + 1 x 16.
        uint32_t va[32], vb[32];
        // ---- pass 1: row maximum over the 197 real keys (3-input max) ----
        float m0 = -INFINITY, m1 = -INFINITY;
        auto max_chunk = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        };
        tmem_ld_32x32(tbuf, va);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 6; ch += 2) {
          tmem_ld_32x32(tbuf + (uint32_t)((ch + 1) * 32), vb);
          max_chunk(va);
          tmem_ld_wait();
          if (ch + 2 < 6) tmem_ld_32x32(tbuf + (uint32_t)((ch + 2) * 32), va);
          else tmem_ld_32x16(tbuf + 192u, va);
          max_chunk(vb);
          tmem_ld_wait();
        }
        m0 = max3(m0, __uint_as_float(va[0]), __uint_as_float(va[1]));
        m1 = max3(m1, __uint_as_float(va[2]), __uint_as_float(va[3]));
        const float m = max3(m0, m1, __uint_as_float(va[4]));        // keys 192..196
        const float2 c2 = make_float2(c, c);
        const float2 nmc2 = make_float2(-m * c, -m * c);
        // ---- pass 2: p = 2^(s c - m c), row sum, P (bf16 pairs) written in place over S ----
        float2 l2 = make_float2(0.f, 0.f);
        auto exp_chunk = [&](const uint32_t (&v)[32], int ch) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 a = ffma2v(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                    c2, nmc2);
            // alternate pairs between the MUFU unit and the FMA-pipe polynomial
            const float2 p = (VPF_ATTN_POLY_EXP2 && (j & 1)) ? ex2_poly2(a) : make_float2(ex2(a.x), ex2(a.y));
            l2 = fadd2v(l2, p);
            pk[j] = pack2(p.x, p.y, f16);
          }
          tmem_st_32x16(tbuf + (uint32_t)(ch * 16), pk);
        };
        tmem_ld_32x32(tbuf, va);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 6; ch += 2) {
          tmem_ld_32x32(tbuf + (uint32_t)((ch + 1) * 32), vb);
          exp_chunk(va, ch);
          tmem_ld_wait();
          if (ch + 2 < 6) tmem_ld_32x32(tbuf + (uint32_t)((ch + 2) * 32), va);
          else tmem_ld_32x16(tbuf + 192u, va);
          exp_chunk(vb, ch + 1);
          tmem_ld_wait();
        }
        {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k0 = 192 + 2 * j;
            const float p0 = (k0 < kT) ? ex2(fmaf(__uint_as_float(va[2 * j]), c, nmc2.x)) : 0.f;
            const float p1 = (k0 + 1 < kT) ? ex2(fmaf(__uint_as_float(va[2 * j + 1]), c, nmc2.x)) : 0.f;
            l2.x += p0; l2.y += p1;
            pk[j] = pack2(p0, p1, f16);
          }
          tmem_st_32x8(tbuf + 96u, pk);
        }
        tmem_st_wait();
        inv_l = 1.0f / (l2.x + l2.y);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      // ---- O = P V is ready: normalise, stage through this unit's Q tile, store coalesced ----
      mbar_wait(&o_full[b], ip & 1, 17);
      tcgen05_fence_after();
      uint8_t* stage = sQ + qs * kQBytes + quarter * 32 * 128;     // this warp's 32 rows
      if (warp_has_rows) {
        uint32_t v[32];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          tmem_ld_32x32(tbuf + (uint32_t)(kOCol + hlf * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack2(__uint_as_float(v[q * 8 + 0]) * inv_l, __uint_as_float(v[q * 8 + 1]) * inv_l, f16);
            o.y = pack2(__uint_as_float(v[q * 8 + 2]) * inv_l, __uint_as_float(v[q * 8 + 3]) * inv_l, f16);
            o.z = pack2(__uint_as_float(v[q * 8 + 4]) * inv_l, __uint_as_float(v[q * 8 + 5]) * inv_l, f16);
            o.w = pack2(__uint_as_float(v[q * 8 + 6]) * inv_l, __uint_as_float(v[q * 8 + 7]) * inv_l, f16);
            const int chunk = hlf * 4 + q;
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = o;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&buf_empty[b]);     // TMEM buffer may take the next unit's S
      if (warp_has_rows) {
        __nv_bfloat16* orow = out + ((size_t)crop * kT + (size_t)(b * 128 + quarter * 32)) * D + head * kD;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3), chunk = lane & 7;
          if (b * 128 + quarter * 32 + r < kT) {
            const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 128 + ((chunk ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(orow + (size_t)r * D + chunk * 8) = val;
          }
        }
      }
      // generic-proxy accesses of the stage are ordered before the TMA (async proxy) refill
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_empty[qs]);       // Q stage may be refilled
      (void)token;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------
// Run-time token count: the same tcgen05 pipeline for 64 < T < 197 (crop sizes 128 .. 208), where
// round 1 fell back to the warp-level mma.sync kernel.  Keys are padded to the next multiple of 16
// (`keys`; the TMA unit zero-fills tokens >= T, the softmax masks them), and a (crop, head) pair has
// one query tile when T <= 128, two otherwise.  A "unit" is one query tile of one pair; units
// alternate between the two TMEM buffers / softmax warpgroups whatever the tile count, so with one
// tile per pair the two warpgroups work on consecutive pairs.  The T = 197 kernel above stays the
// specialised (fully unrolled) version of this one.
// With ONE query tile per pair (T <= 128) there are two MMA-issuing threads, one per TMEM buffer
// (warps 1 and 10), each walking its own chain S -> PV -> S ... with blocking waits: a single issuer
// that walks both chains in a fixed order waits for buffer 0's O to be read out while buffer 1's PV
// is ready.  Measured (N = 4096, 12 heads): T = 101 0.762 -> 0.625 ms, T = 65 0.665 -> 0.531 ms.
// With TWO tiles per pair the same split is SLOWER (T = 145 1.10 -> 1.28 ms, T = 170 1.31 -> 1.47 ms,
// the T = 197 kernel 1.087 -> 1.32 ms, with or without an anti-phase start), so those keep the single
// issuer (experiments/README.md).
//   TMEM buffer b at column 256 b:  S [0,keys)  P aliases [0,keys/2)  O [128,192)
__global__ void __launch_bounds__(kThreadsVar, 1)
attention_tc_var_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        __nv_bfloat16* __restrict__ out, int n_pairs, int heads, int D, int f16_,
                        int T, int keys) {
  const bool f16 = f16_ != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = base;                                   // [kQStages][128][128 B]
  uint8_t* sK = sQ + kQStages * kQBytes;                // [kKVStages][<= 208][128 B]
  uint8_t* sV = sK + kKVStages * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kKVBytes);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 4;
  uint64_t* kv_full = bars + 8;
  uint64_t* kv_empty = bars + 10;
  uint64_t* s_full = bars + 12;
  uint64_t* p_full = bars + 14;
  uint64_t* o_full = bars + 16;
  uint64_t* buf_empty = bars + 18;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_pairs = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tpp = T > 128 ? 2 : 1;                       // query tiles per pair
  const int units = my_pairs * tpp;
  const uint32_t kv_bytes = (uint32_t)keys * 128u;       // K (or V) tile: keys rows x 128 B

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < kQStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 4); }
    for (int i = 0; i < kKVStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&buf_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int u = 0; u < units; ++u) {
        const int ip = u / tpp, t = u - ip * tpp;
        const int pair = (int)blockIdx.x + ip * (int)gridDim.x;
        const int crop = pair / heads, head = pair - crop * heads;
        if (t == 0) {
          const int ks = ip & 1;
          mbar_wait(&kv_empty[ks], ((ip >> 1) & 1) ^ 1, 10);
          mbar_arrive_expect_tx(&kv_full[ks], 2 * kv_bytes);
          tma_load_3d(sK + ks * kKVBytes, &tmKV, &kv_full[ks], D + head * kD, 0, crop);
          tma_load_3d(sV + ks * kKVBytes, &tmKV, &kv_full[ks], 2 * D + head * kD, 0, crop);
        }
        const int qs = u & 3;
        mbar_wait(&q_empty[qs], ((u >> 2) & 1) ^ 1, 11);
        mbar_arrive_expect_tx(&q_full[qs], kQBytes);
        tma_load_3d(sQ + qs * kQBytes, &tmQ, &q_full[qs], head * kD, t * 128, crop);
      }
    }
  } else if (warp == 1 || warp == kMmaWarp1) {
    // ===================== MMA issuers =====================
    // One query tile per pair (T <= 128): warp 1 walks TMEM buffer 0's chain (even units), warp 10
    // buffer 1's, each S -> PV -> S ... with blocking waits.  Two tiles per pair: warp 1 alone in
    // the fixed order PV(u), S(u+2), PV(u+1), ... (two issuers are slower there, see above).
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_16(128, keys, f16);
      const uint32_t idesc_o = make_idesc_bmn(128, kD, f16);
      auto issue_s = [&](int u) {
        const int ip = u / tpp, b = u & 1, qs = u & 3, ks = ip & 1, k = u >> 1;
        mbar_wait(&q_full[qs], (u >> 2) & 1, 12);
        mbar_wait(&kv_full[ks], (ip >> 1) & 1, 13);      // already complete for a pair's 2nd tile
        mbar_wait(&buf_empty[b], (k & 1) ^ 1, 14);
        tcgen05_fence_after();
        const uint64_t adesc = make_smem_desc_k_sw128(smem_u32(sQ + qs * kQBytes));
        const uint64_t bdesc = make_smem_desc_k_sw128(smem_u32(sK + ks * kKVBytes));
        const uint32_t d = tmem_base + (uint32_t)(b * kBufCols);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_bf16(d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc_s, (uint32_t)(kk != 0));
        umma_commit(&s_full[b]);
      };
      auto issue_pv = [&](int u) {
        const int ip = u / tpp, t = u - ip * tpp, b = u & 1, ks = ip & 1, k = u >> 1;
        mbar_wait(&p_full[b], k & 1, 15);
        tcgen05_fence_after();
        const uint32_t vaddr = smem_u32(sV + ks * kKVBytes);
        const uint32_t p = tmem_base + (uint32_t)(b * kBufCols);
        const uint32_t d = p + (uint32_t)kOCol;
        const int groups = keys >> 4;
#pragma unroll 1
        for (int kk = 0; kk < groups; ++kk)              // 16 keys: 8 packed P columns, 2 V row atoms
          umma_bf16_ts(d, p + (uint32_t)(kk * 8), make_smem_desc_mn_sw128(vaddr + kk * 2048), idesc_o,
                       (uint32_t)(kk != 0));
        umma_commit(&o_full[b]);
        if (t == tpp - 1) umma_commit(&kv_empty[ks]);    // every MMA of this pair has retired
      };
      if (tpp == 1) {
        for (int u = (warp == 1 ? 0 : 1); u < units; u += 2) {
          issue_s(u);
          issue_pv(u);
        }
      } else if (warp == 1) {
        if (units > 0) issue_s(0);
        if (units > 1) issue_s(1);
        for (int u = 0; u < units; ++u) {
          issue_pv(u);
          if (u + 2 < units) issue_s(u + 2);
        }
      }
    }
  } else {
    // ===================== softmax / output warpgroups =====================
    const int b = (warp - 2) >> 2;                 // TMEM buffer == unit parity
    const int quarter = warp & 3;                  // TMEM lane quarter
    const int t = tpp == 2 ? b : 0;                // query tile this warpgroup works on
    const bool warp_has_rows = (t * 128 + quarter * 32) < T;
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t tbuf = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * kBufCols);
    const int nchunks = (keys + 31) >> 5;          // 32-column chunks; the last one may be half real
    for (int u = b; u < units; u += 2) {
      const int ip = u / tpp, k = u >> 1, qs = u & 3;
      const int pair = (int)blockIdx.x + ip * (int)gridDim.x;
      const int crop = pair / heads, head = pair - crop * heads;
      mbar_wait(&s_full[b], k & 1, 16);
      tcgen05_fence_after();
      float inv_l = 0.f;
      if (warp_has_rows) {
        uint32_t va[32], vb[32];
        // ---- pass 1: row maximum over the T real keys ----
        float m0 = -INFINITY, m1 = -INFINITY;
        auto max_chunk = [&](const uint32_t (&v)[32], int ch) {
          if (ch + 1 < nchunks) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          } else {                                 // last chunk: columns >= T are padding
            const int lim = T - ch * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < lim) m0 = fmaxf(m0, __uint_as_float(v[j]));
          }
        };
        int ch = 0;
        tmem_ld_32x32(tbuf, va);
        tmem_ld_wait();
        while (true) {
          if (ch + 1 < nchunks) tmem_ld_32x32(tbuf + (uint32_t)((ch + 1) * 32), vb);
          max_chunk(va, ch);
          if (ch + 1 >= nchunks) break;
          tmem_ld_wait();
          if (ch + 2 < nchunks) tmem_ld_32x32(tbuf + (uint32_t)((ch + 2) * 32), va);
          max_chunk(vb, ch + 1);
          if (ch + 2 >= nchunks) break;
          tmem_ld_wait();
          ch += 2;
        }
        const float m = fmaxf(m0, m1);
        const float2 c2 = make_float2(c, c);
        const float nmc = -m * c;
        const float2 nmc2 = make_float2(nmc, nmc);
        // ---- pass 2: p = 2^(s c - m c), row sum, P (bf16 pairs) written in place over S ----
        float2 l2 = make_float2(0.f, 0.f);
        auto exp_chunk = [&](const uint32_t (&v)[32], int chn) {
          uint32_t pk[16];
          if (chn + 1 < nchunks) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 a = ffma2v(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                      c2, nmc2);
              const float2 p = (VPF_ATTN_POLY_EXP2 && (j & 1)) ? ex2_poly2(a) : make_float2(ex2(a.x), ex2(a.y));
              l2 = fadd2v(l2, p);
              pk[j] = pack2(p.x, p.y, f16);
            }
          } else {
            const int lim = T - chn * 32;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float p0 = (2 * j < lim) ? ex2(fmaf(__uint_as_float(v[2 * j]), c, nmc)) : 0.f;
              const float p1 = (2 * j + 1 < lim) ? ex2(fmaf(__uint_as_float(v[2 * j + 1]), c, nmc)) : 0.f;
              l2.x += p0; l2.y += p1;
              pk[j] = pack2(p0, p1, f16);
            }
          }
          tmem_st_32x16(tbuf + (uint32_t)(chn * 16), pk);
        };
        ch = 0;
        tmem_ld_32x32(tbuf, va);
        tmem_ld_wait();
        while (true) {
          if (ch + 1 < nchunks) tmem_ld_32x32(tbuf + (uint32_t)((ch + 1) * 32), vb);
          exp_chunk(va, ch);
          if (ch + 1 >= nchunks) break;
          tmem_ld_wait();
          if (ch + 2 < nchunks) tmem_ld_32x32(tbuf + (uint32_t)((ch + 2) * 32), va);
          exp_chunk(vb, ch + 1);
          if (ch + 2 >= nchunks) break;
          tmem_ld_wait();
          ch += 2;
        }
        tmem_st_wait();
        inv_l = 1.0f / (l2.x + l2.y);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      // ---- O = P V is ready: normalise, stage through this unit's Q tile, store coalesced ----
      mbar_wait(&o_full[b], k & 1, 17);
      tcgen05_fence_after();
      uint8_t* stage = sQ + qs * kQBytes + quarter * 32 * 128;     // this warp's 32 rows
      if (warp_has_rows) {
        uint32_t v[32];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          tmem_ld_32x32(tbuf + (uint32_t)(kOCol + hlf * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack2(__uint_as_float(v[q * 8 + 0]) * inv_l, __uint_as_float(v[q * 8 + 1]) * inv_l, f16);
            o.y = pack2(__uint_as_float(v[q * 8 + 2]) * inv_l, __uint_as_float(v[q * 8 + 3]) * inv_l, f16);
            o.z = pack2(__uint_as_float(v[q * 8 + 4]) * inv_l, __uint_as_float(v[q * 8 + 5]) * inv_l, f16);
            o.w = pack2(__uint_as_float(v[q * 8 + 6]) * inv_l, __uint_as_float(v[q * 8 + 7]) * inv_l, f16);
            const int chunk = hlf * 4 + q;
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = o;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&buf_empty[b]);     // TMEM buffer may take the next unit's S
      if (warp_has_rows) {
        __nv_bfloat16* orow = out + ((size_t)crop * T + (size_t)(t * 128 + quarter * 32)) * D + head * kD;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3), chunk = lane & 7;
          if (t * 128 + quarter * 32 + r < T) {
            const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 128 + ((chunk ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(orow + (size_t)r * D + chunk * 8) = val;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_empty[qs]);       // Q stage may be refilled
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace attn
}  // namespace vpf
