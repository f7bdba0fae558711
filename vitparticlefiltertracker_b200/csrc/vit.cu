// K2: timm-style ViT/16 encoder on sm_100a.  Oracle: oracle/vit.py (SURVEY.md Appendix A D1;
// upstream anchor README.md:7 "Vision Transformer (ViT) ... feature extraction").
//
// Per block:  LN1 -> QKV GEMM(+bias) -> fused attention -> proj GEMM(+bias+residual)
//             LN2 -> FC1 GEMM(+bias+GELU) -> FC2 GEMM(+bias+residual)
// Residual stream X is fp32 [n*197][D]; GEMM operands are bf16; accumulation fp32 in TMEM.
#include "common.cuh"
#include "gemm.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"

#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace vpf {

// ---------------------------------------------------------------------------------------------
// error string (thread local)
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_for_rows(long long rows) {
  static int v = -2;                       // -1 auto, 0 off, 1 on
  if (v == -2) {
    const char* e = getenv("VPF_PDL");
    v = !e ? -1 : (e[0] == '0' ? 0 : 1);
  }
  return v < 0 ? rows <= kPdlMaxRows : v == 1;
}

// ---------------------------------------------------------------------------------------------
// TMA descriptors: the driver entry point is resolved at run time (libvpf.so links only the
// static CUDA runtime, so it loads on machines without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// bf16 row-major [rows][cols] matrix, tile = box_rows x 64 columns, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols,
                    uint32_t box_rows, uint64_t pitch_elems = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return VPF_ERR_CUDA; }
  if (!pitch_elems) pitch_elems = cols;
  if (((uintptr_t)ptr & 15) || (pitch_elems * 2) % 16) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    return VPF_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return VPF_ERR_CUDA; }
  return VPF_OK;
}

// Row-major [rows][cols] output tile map for the epilogue's TMA traffic, 32 x 32 tiles:
// bf16 (64-byte rows, 64-byte swizzle) or fp32 (128-byte rows, 128-byte swizzle).
static int make_store_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols,
                          bool f32 = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return VPF_ERR_CUDA; }
  const uint64_t esz = f32 ? 4 : 2;
  if (((uintptr_t)ptr & 15) || (cols * esz) % 16) {
    set_error("TMA store target must be 16-byte aligned with a 16-byte multiple row pitch");
    return VPF_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * esz};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(store) failed: %d", (int)r); return VPF_ERR_CUDA; }
  return VPF_OK;
}

// The epilogue's tensor map depends only on (pointer, rows, cols, element type): encoded once and
// kept -- a frame re-uses the same ~10 output buffers for its ~60 GEMM launches, and each encode is
// a driver call on the host's critical path (visible at ViT-Ti scale, where a frame is 2 ms).
static int cached_store_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, bool f32) {
  struct Entry { const void* ptr; uint64_t rows, cols; bool f32; CUtensorMap map; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
      if (e.ptr == ptr && e.rows == rows && e.cols == cols && e.f32 == f32) { *map = e.map; return VPF_OK; }
  }
  int rc = make_store_map(map, ptr, rows, cols, f32);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() >= 256) cache.clear();
  cache.push_back(Entry{ptr, rows, cols, f32, *map});
  return VPF_OK;
}

static int num_sms();
static int num_sms_any() { return num_sms(); }

static int pick_block_n(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  return 0;
}

// 3-D view [crop][token][3D] of the fused qkv buffer: the TMA unit zero-fills tokens >= T.
static int make_qkv_map(CUtensorMap* map, const void* ptr, int n_crops, int D, uint32_t box_tokens,
                        int T, int ncols3 = 3) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return VPF_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)ncols3 * D, (cuuint64_t)T, (cuuint64_t)n_crops};
  cuuint64_t strides[2] = {(cuuint64_t)ncols3 * D * 2, (cuuint64_t)T * ncols3 * D * 2};
  cuuint32_t box[3] = {64, box_tokens, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(qkv) failed: %d", (int)r); return VPF_ERR_CUDA; }
  return VPF_OK;
}

// per-device "already done" flag for per-kernel one-time setup (function attributes are per context)
static bool first_time_on_device(bool (&done)[32]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 31;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// SM count of the CURRENT device (cached per device; a benign race: every writer stores the
// same value).  CTA pairs need an even number of CTAs.
static int num_sms() {
  static int cache[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = cache[dev & 31];
  if (!v) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    v = n & ~1;
  }
  return v;
}

// VPF_GEMM_CTAS=1 selects the single-CTA (cta_group::1) kernel; default is the CTA-pair kernel.
static int gemm_ctas() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("VPF_GEMM_CTAS");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

template <int BLOCK_N, int MODE, int CTAS>
static int launch_gemm_c(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g,
                         cudaStream_t s) {
  using Cfg = GemmCfg<BLOCK_N, CTAS, MODE>;
  auto kern = gemm_bf16_kernel<BLOCK_N, MODE, CTAS>;
  static bool attr_set[32] = {};
  if (first_time_on_device(attr_set))
    VPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
  const int tiles = ((g.M + kBlockM * CTAS - 1) / (kBlockM * CTAS)) * (g.N / BLOCK_N);
  const int max_groups = num_sms() / CTAS;
  const int groups = tiles < max_groups ? tiles : max_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CTAS);
  cfg.blockDim = dim3(EpiCfg<BLOCK_N, MODE>::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_for_rows(g.M) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CUtensorMap tmC = tmA;                       // modes 2 / 3 do not use it
  if (mode_out16(MODE) || MODE == kBiasResidualF32Tma) {
    int rc = cached_store_map(&tmC, g.out, (uint64_t)g.M, (uint64_t)g.N,
                              MODE == kBiasResidualF32Tma && !g.x16);
    if (rc) return rc;
  }
  VPF_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, g));
  return VPF_OK;
}

template <int BLOCK_N, int MODE>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g,
                         cudaStream_t s) {
  return gemm_ctas() == 2 ? launch_gemm_c<BLOCK_N, MODE, 2>(tmA, tmB, g, s)
                          : launch_gemm_c<BLOCK_N, MODE, 1>(tmA, tmB, g, s);
}

template <int BLOCK_N>
static int launch_gemm_n(const CUtensorMap& a, const CUtensorMap& b, const GemmArgs& g, int mode,
                         cudaStream_t s) {
  switch (mode) {
    case kBiasBf16: return launch_gemm_t<BLOCK_N, kBiasBf16>(a, b, g, s);
    case kBiasGeluBf16: return launch_gemm_t<BLOCK_N, kBiasGeluBf16>(a, b, g, s);
    case kBiasResidualF32:
      // small K: the GEMM is bound by the fp32 read-modify-write of the residual -> TMA epilogue
      // (two tiles in flight per warp); large K: MMA-bound, keep the deeper smem pipeline
      return g.K <= 1536 ? launch_gemm_t<BLOCK_N, kBiasResidualF32Tma>(a, b, g, s)
                         : launch_gemm_t<BLOCK_N, kBiasResidualF32>(a, b, g, s);
    case kPatchEmbedF32: return launch_gemm_t<BLOCK_N, kPatchEmbedF32>(a, b, g, s);
    case kLnBiasBf16: return launch_gemm_t<BLOCK_N, kLnBiasBf16>(a, b, g, s);
    case kLnBiasGeluBf16: return launch_gemm_t<BLOCK_N, kLnBiasGeluBf16>(a, b, g, s);
  }
  set_error("gemm: unknown epilogue mode %d", mode);
  return VPF_ERR_INVALID;
}

static int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const GemmArgs& g, int mode,
                       cudaStream_t s) {
  switch (pick_block_n(g.N)) {
    case 256: return launch_gemm_n<256>(a, b, g, mode, s);
    case 192: return launch_gemm_n<192>(a, b, g, mode, s);
    case 128: return launch_gemm_n<128>(a, b, g, mode, s);
  }
  set_error("gemm: N=%d must be a multiple of 128", g.N);
  return VPF_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (D <= 1024), fp32 statistics.
template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x,
                                                         size_t row_stride,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta,
                                                         OutT* __restrict__ out, int rows, int D,
                                                         bool f16 = false, bool x16 = false) {
  pdl_launch_dependents();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = D >> 2;
  // x16: the residual stream is fp16 (`x` then points at halves; row_stride still in elements)
  const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)row * row_stride);
  const uint2* h4 = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x) +
                                                   (size_t)row * row_stride);
  pdl_wait();
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      if (x16) {
        const uint2 h = h4[idx];
        const float2 lo = unpack_h2(h.x), hi = unpack_h2(h.y);
        v[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        v[i] = x4[idx];
      }
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + 1e-6f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
      float o0 = (v[i].x - mean) * rstd * g.x + b.x;
      float o1 = (v[i].y - mean) * rstd * g.y + b.y;
      float o2 = (v[i].z - mean) * rstd * g.z + b.z;
      float o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (sizeof(OutT) == 2) {
        uint2 p;
        p.x = pack16x2(o0, o1, f16);
        p.y = pack16x2(o2, o3, f16);
        reinterpret_cast<uint2*>(out + (size_t)row * D)[idx] = p;
      } else {
        reinterpret_cast<float4*>(out + (size_t)row * D)[idx] = make_float4(o0, o1, o2, o3);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm of the fp16 residual stream (the per-block hot case: 22 launches per ViT-B frame).
// In-step timing (bench.py encoder_breakdown_ms, round 2) showed the one-row-per-warp kernel above
// taking 1.07 ms per launch inside the power-capped step against 0.37 ms alone: the encoder's GEMMs
// hold the SM clock near 1.2 GHz, and at that clock the kernel is bound by instruction issue and by
// the churn of 100 000 short-lived blocks, not by HBM.  This version is persistent (each warp walks
// rows with a grid stride), loads 16-byte vectors (8 halves), issues the next row's loads before
// it reduces the current one, keeps gamma / beta in shared memory, and does its arithmetic as
// packed fp32x2 (FADD2 / FFMA2): ~130 instead of ~240 instructions per row and lane.  Same
// arithmetic as above (fp32 mean, centred sum of squares, eps 1e-6), so results are unchanged up to
// the fp32 rounding of the packed operations.
template <int NV>                           // 16-byte vectors per lane: D <= NV * 256
__global__ void __launch_bounds__(256, (NV <= 3 ? 4 : 3)) layernorm16_kernel(const __half* __restrict__ x,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           uint16_t* __restrict__ out, int rows,
                                                           int D, bool f16) {
  extern __shared__ __align__(16) float s_gb[];          // [D] gamma | [D] beta
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_gb[i] = gamma[i]; s_gb[D + i] = beta[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float inv_d = 1.0f / (float)D;
  pdl_wait();
  // one row in registers ahead of the one being reduced, and the rows after it pulled into L2
  // with prefetch instructions (no registers): HBM latency is covered three rows deep per warp
  uint4 raw[NV];
  auto load_row = [&](int r) {
    const uint4* p = reinterpret_cast<const uint4*>(x + (size_t)r * D);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) raw[i] = p[lane + 32 * i];
  };
  auto prefetch_row = [&](int r) {                       // one 128-byte line per lane
    if (r < rows && lane * 64 < D)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (size_t)r * D + lane * 64));
  };
  if (row < rows) load_row(row);
  prefetch_row(row + warps);
  prefetch_row(row + 2 * warps);
  for (; row < rows; row += warps) {
    float2 v[NV][4];
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        v[i][0] = unpack_h2(raw[i].x); v[i][1] = unpack_h2(raw[i].y);
        v[i][2] = unpack_h2(raw[i].z); v[i][3] = unpack_h2(raw[i].w);
        acc = fadd2(acc, fadd2(fadd2(v[i][0], v[i][1]), fadd2(v[i][2], v[i][3])));
      }
    }
    if (row + warps < rows) load_row(row + warps);        // next row in flight during the reductions
    prefetch_row(row + 3 * warps);
    const float mean = warp_sum(acc.x + acc.y) * inv_d;
    const float2 nmean = make_float2(-mean, -mean);
    float2 sq = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[i][k] = fadd2(v[i][k], nmean);
          sq = ffma2(v[i][k], v[i][k], sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq.x + sq.y) * inv_d + 1e-6f);
    const float2 rstd2 = make_float2(rstd, rstd);
    const float2 zero2 = make_float2(0.f, 0.f);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = *reinterpret_cast<const float4*>(s_gb + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(s_gb + vi * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_gb + D + vi * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(s_gb + D + vi * 8 + 4);
        const float2 y0 = ffma2(ffma2(v[i][0], rstd2, zero2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
        const float2 y1 = ffma2(ffma2(v[i][1], rstd2, zero2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
        const float2 y2 = ffma2(ffma2(v[i][2], rstd2, zero2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
        const float2 y3 = ffma2(ffma2(v[i][3], rstd2, zero2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
        uint4 pk;
        pk.x = pack16x2(y0.x, y0.y, f16); pk.y = pack16x2(y1.x, y1.y, f16);
        pk.z = pack16x2(y2.x, y2.y, f16); pk.w = pack16x2(y3.x, y3.y, f16);
        o[vi] = pk;
      }
    }
  }
}

template <int NV>
static cudaError_t launch_layernorm16_nv(cudaStream_t s, const __half* x, const float* gamma,
                                         const float* beta, uint16_t* out, int rows, int D, bool f16) {
  // persistent: exactly as many blocks as are resident at once (a later wave would start its
  // grid-stride share late), never more warps than rows
  static int per_sm[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& occ = per_sm[dev & 31];
  if (!occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm16_kernel<NV>, 256,
                                                      (size_t)2 * 1024 * sizeof(float)) != cudaSuccess || occ < 1)
      occ = 2;
  }
  int blocks = num_sms_any() * occ;
  const int need = (rows + 7) / 8;
  if (blocks > need) blocks = need;
  return launch_pdl(pdl_for_rows(rows), layernorm16_kernel<NV>, dim3(blocks), dim3(256),
                    (size_t)2 * D * sizeof(float), s, x, gamma, beta, out, rows, D, f16);
}

template <typename OutT>
static cudaError_t launch_layernorm(int blocks, cudaStream_t s, const float* x, size_t row_stride,
                                    const float* gamma, const float* beta, OutT* out, int rows,
                                    int D, bool f16 = false, bool x16 = false) {
  // the hot case: contiguous fp16 stream, 16-bit output, D a multiple of 8 up to 1024
  // (-DVPF_LN_LEGACY=1 builds the A/B variant that keeps the one-row-per-warp kernel everywhere)
#ifndef VPF_LN_LEGACY
#define VPF_LN_LEGACY 0
#endif
  if (!VPF_LN_LEGACY && sizeof(OutT) == 2 && x16 && row_stride == (size_t)D && D % 8 == 0 && D <= 1024) {
    const __half* xh = reinterpret_cast<const __half*>(x);
    uint16_t* o = reinterpret_cast<uint16_t*>(out);
    if (D <= 256) return launch_layernorm16_nv<1>(s, xh, gamma, beta, o, rows, D, f16);
    if (D <= 512) return launch_layernorm16_nv<2>(s, xh, gamma, beta, o, rows, D, f16);
    if (D <= 768) return launch_layernorm16_nv<3>(s, xh, gamma, beta, o, rows, D, f16);
    return launch_layernorm16_nv<4>(s, xh, gamma, beta, o, rows, D, f16);
  }
  return launch_pdl(pdl_for_rows(rows), layernorm_kernel<OutT>, dim3(blocks), dim3(256), 0, s, x, row_stride, gamma, beta,
                    out, rows, D, f16, x16);
}

// ---------------------------------------------------------------------------------------------
// Row statistics of the fp16 residual stream for the LayerNorm-folded GEMMs: stats[row] = (mean,
// 1/sqrt(var + 1e-6)), same arithmetic as layernorm16_kernel (fp32 mean, centred sum of squares).
// Reads the stream once (2 B/element) and writes 8 B per row; used where no residual GEMM precedes
// the LayerNorm (after the patch embedding) and as the A/B reference of the epilogue-fused sums.
template <int NV>
__global__ void __launch_bounds__(256, 4) rowstats16_kernel(const __half* __restrict__ x,
                                                            float2* __restrict__ stats, int rows, int D) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float inv_d = 1.0f / (float)D;
  pdl_wait();
  uint4 raw[NV];
  auto load_row = [&](int r) {
    const uint4* p = reinterpret_cast<const uint4*>(x + (size_t)r * D);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) raw[i] = p[lane + 32 * i];
  };
  if (row < rows) load_row(row);
  for (; row < rows; row += warps) {
    float2 v[NV][4];
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        v[i][0] = unpack_h2(raw[i].x); v[i][1] = unpack_h2(raw[i].y);
        v[i][2] = unpack_h2(raw[i].z); v[i][3] = unpack_h2(raw[i].w);
        acc = fadd2(acc, fadd2(fadd2(v[i][0], v[i][1]), fadd2(v[i][2], v[i][3])));
      }
    }
    if (row + warps < rows) load_row(row + warps);
    const float mean = warp_sum(acc.x + acc.y) * inv_d;
    const float2 nmean = make_float2(-mean, -mean);
    float2 sq = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 c = fadd2(v[i][k], nmean);
          sq = ffma2(c, c, sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq.x + sq.y) * inv_d + 1e-6f);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
  }
}

template <int NV>
static cudaError_t launch_rowstats16_nv(cudaStream_t s, const __half* x, float2* stats, int rows, int D) {
  int blocks = num_sms_any() * 4;
  const int need = (rows + 7) / 8;
  if (blocks > need) blocks = need;
  return launch_pdl(pdl_for_rows(rows), rowstats16_kernel<NV>, dim3(blocks), dim3(256), 0, s, x, stats,
                    rows, D);
}
static cudaError_t launch_rowstats16(cudaStream_t s, const void* x16, float2* stats, int rows, int D) {
  const __half* x = reinterpret_cast<const __half*>(x16);
  if (D <= 256) return launch_rowstats16_nv<1>(s, x, stats, rows, D);
  if (D <= 512) return launch_rowstats16_nv<2>(s, x, stats, rows, D);
  if (D <= 768) return launch_rowstats16_nv<3>(s, x, stats, rows, D);
  return launch_rowstats16_nv<4>(s, x, stats, rows, D);
}

// column groups a residual GEMM with N = D splits a row into (its n-blocks x 2 epilogue halves)
static int ln_stats_parts(int D) { const int bn = pick_block_n(D); return bn ? (D / bn) * 2 : 0; }

// X[crop*T][:] = cls + pos[0]
__global__ void cls_init_kernel(float* __restrict__ X, const float* __restrict__ cls,
                                const float* __restrict__ pos, int n_crops, int D, int T, bool x16) {
  pdl_launch_dependents();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_crops * D) return;
  pdl_wait();       // X is still read by the previous forward's last kernels
  int crop = i / D, d = i - crop * D;
  const float v = cls[d] + pos[d];
  if (x16) reinterpret_cast<__half*>(X)[(size_t)crop * T * D + d] = __float2half_rn(v);
  else X[(size_t)crop * T * D + d] = v;
}

// VPF_ATTN=mma selects the warp-level mma.sync kernel; default is the tcgen05 kernel.
static bool attn_use_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VPF_ATTN");
    v = (e && e[0] == 'm') ? 0 : 1;
  }
  return v == 1;
}

template <int KT>
static int launch_attention_small(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_crops,
                                  int heads, int tokens, cudaStream_t s, int f16) {
  static bool attr[32] = {};
  if (first_time_on_device(attr))
    VPF_CUDA(cudaFuncSetAttribute(attention_small_kernel<KT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SmallAttn<KT>::kSmem));
  const int n_pairs = n_crops * heads;
  VPF_CUDA(launch_pdl(pdl_for_rows((long long)n_crops * tokens), attention_small_kernel<KT>, dim3((n_pairs + kSmallWarps - 1) / kSmallWarps),
                      dim3(kSmallWarps * 32), (size_t)SmallAttn<KT>::kSmem, s, qkv, out, n_pairs,
                      heads, tokens, f16 != 0));
  return VPF_OK;
}

// T = 197 (224-pixel crops): the specialised tcgen05 kernel; 64 < T < 197 (crop sizes 128 .. 208):
// the same pipeline with a run-time token count; T <= 64: one warp per (crop, head).  The
// CTA-per-pair mma.sync kernel remains only behind VPF_ATTN=mma (A/B runs).
static int keys_padded(int tokens) { return (tokens + 15) & ~15; }

static int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_crops, int heads,
                            int tokens, const CUtensorMap* mq, const CUtensorMap* mkv,
                            cudaStream_t s, int f16 = 0) {
  const int D = heads * kHeadDim;
  if (tokens <= 32) return launch_attention_small<32>(qkv, out, n_crops, heads, tokens, s, f16);
  if (tokens <= 64) return launch_attention_small<64>(qkv, out, n_crops, heads, tokens, s, f16);
  if (attn_use_tc()) {
    const int n_pairs = n_crops * heads;
    const int grid = n_pairs < num_sms() ? n_pairs : num_sms();
    const bool pdl = pdl_for_rows((long long)n_crops * tokens);
    if (tokens == attn::kT) {
      static bool attr[32] = {};
      if (first_time_on_device(attr))
        VPF_CUDA(cudaFuncSetAttribute(attn::attention_tc_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, attn::kSmemBytes));
      VPF_CUDA(launch_pdl(pdl, attn::attention_tc_kernel, dim3(grid), dim3(attn::kThreads),
                          (size_t)attn::kSmemBytes, s, *mq, *mkv, out, n_pairs, heads, D, f16));
    } else {
      static bool attr[32] = {};
      if (first_time_on_device(attr))
        VPF_CUDA(cudaFuncSetAttribute(attn::attention_tc_var_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, attn::kSmemBytes));
      VPF_CUDA(launch_pdl(pdl, attn::attention_tc_var_kernel, dim3(grid), dim3(attn::kThreadsVar),
                          (size_t)attn::kSmemBytes, s, *mq, *mkv, out, n_pairs, heads, D, f16, tokens,
                          keys_padded(tokens)));
    }
  } else {
    static bool attr[32] = {};
    if (first_time_on_device(attr))
      VPF_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kAttnSmem));
    VPF_CUDA(launch_pdl(pdl_for_rows((long long)n_crops * tokens), attention_kernel, dim3(heads * n_crops), dim3(kAttnThreads),
                        (size_t)kAttnSmem, s, qkv, out, D, tokens, f16 != 0));
  }
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

// ---- CLS-only last block (exact: the encoder output is the CLS row, and in the last block a
// token's output depends on the other tokens only through their keys and values) --------------
// Xc[crop][:] = X[crop*197][:]
__global__ void gather_cls_kernel(const float* __restrict__ X, float* __restrict__ Xc, int n_crops,
                                  int D, int T, bool x16) {
  pdl_launch_dependents();
  int i = blockIdx.x * blockDim.x + threadIdx.x;     // float4 index
  const int nv = D >> 2;
  if (i >= n_crops * nv) return;
  pdl_wait();
  int crop = i / nv, v = i - crop * nv;
  if (x16) {
    const uint2 h = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(X) +
                                                   (size_t)crop * T * D)[v];
    const float2 lo = unpack_h2(h.x), hi = unpack_h2(h.y);
    reinterpret_cast<float4*>(Xc)[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
  } else {
    reinterpret_cast<float4*>(Xc)[i] =
        reinterpret_cast<const float4*>(X + (size_t)crop * T * D)[v];
  }
}

// One warp per (crop, head): o = softmax(q K^T / 8) V for the single CLS query.
// q: bf16 [n_crops][D]; kv: bf16 [n_crops*T][2D] (K then V); o: bf16 [n_crops][D].  T <= 224.
__global__ void __launch_bounds__(256) cls_attention_kernel(const __nv_bfloat16* __restrict__ q,
                                                            const __nv_bfloat16* __restrict__ kv,
                                                            __nv_bfloat16* __restrict__ o,
                                                            int n_pairs, int heads, bool f16,
                                                            int T) {
  pdl_launch_dependents();
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const int crop = pair / heads, head = pair - crop * heads;
  const int D = heads * kHeadDim;
  pdl_wait();
  float qf[kHeadDim];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q + (size_t)crop * D + head * kHeadDim);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 t = __ldg(q4 + i);
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack16x2(h[j], f16);
        qf[i * 8 + j * 2] = f.x; qf[i * 8 + j * 2 + 1] = f.y;
      }
    }
  }
  const __nv_bfloat16* kbase = kv + (size_t)crop * T * 2 * D + head * kHeadDim;
  const float c = 0.125f * 1.4426950408889634f;
  float sc[7];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int j = i * 32 + lane;
    float acc = -INFINITY;
    if (j < T) {
      const uint4* k4 = reinterpret_cast<const uint4*>(kbase + (size_t)j * 2 * D);
      acc = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 t = __ldg(k4 + u);
        const uint32_t* h = reinterpret_cast<const uint32_t*>(&t);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          float2 f = unpack16x2(h[w], f16);
          acc = fmaf(qf[u * 8 + w * 2], f.x, acc);
          acc = fmaf(qf[u * 8 + w * 2 + 1], f.y, acc);
        }
      }
      acc *= c;
    }
    sc[i] = acc;
    m = fmaxf(m, acc);
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    sc[i] = exp2f(sc[i] - m);          // exp2(-inf) = 0 for the padded keys
    l += sc[i];
  }
  l = warp_sum(l);
  const __nv_bfloat16* vbase = kbase + D + 2 * lane;
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const int j = i * 32 + jj;
      const float p = __shfl_sync(0xffffffffu, sc[i], jj);
      if (j < T) {
        float2 f = unpack16x2(*reinterpret_cast<const uint32_t*>(vbase + (size_t)j * 2 * D), f16);
        o0 = fmaf(p, f.x, o0);
        o1 = fmaf(p, f.y, o1);
      }
    }
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<uint32_t*>(o + (size_t)crop * D + head * kHeadDim + 2 * lane) =
      pack16x2(o0 * inv, o1 * inv, f16);
}

struct LayerW {
  const float *ln1_g, *ln1_b, *qkv_b, *proj_b, *ln2_g, *ln2_b, *fc1_b, *fc2_b;
  const void *qkv_w, *proj_w, *fc1_w, *fc2_w;
  CUtensorMap m_qkv, m_proj, m_fc1, m_fc2;
  // LayerNorm folded into qkv / fc1 (vpf_vit_set_ln_fold): W' = centred gamma (.) W, b' = b + W beta
  const void *qkv_wf = nullptr, *fc1_wf = nullptr;
  const float *qkv_bf = nullptr, *fc1_bf = nullptr;
  CUtensorMap m_qkv_f, m_fc1_f;
};

}  // namespace vpf

using namespace vpf;

struct vpf_vit {
  vpf_vit_config cfg;
  bool x16 = false;             // residual stream X in fp16 (cfg.residual_dtype), else fp32
  int tokens = kMaxTokens;      // T = ppc + 1
  int ppc = kMaxTokens - 1;     // patches per crop = (crop_size / 16)^2
  bool has_weights = false, has_ws = false;
  const void* patch_w = nullptr;
  const float *patch_b = nullptr, *cls = nullptr, *pos = nullptr, *norm_g = nullptr,
              *norm_b = nullptr;
  CUtensorMap m_patch;
  std::vector<LayerW> layers;
  // workspace
  float* X = nullptr;           // [max_rows][D] fp32 residual stream
  __nv_bfloat16* Xn = nullptr;  // [max_rows][D] LN output / attention output
  __nv_bfloat16* big = nullptr; // [max_rows][4D] qkv ([..][3D]) or fc1 output
  uint8_t* cbuf = nullptr;      // compact operands of the CLS-only last block (cls_bytes)
  // activation tensor maps, cached for the last n_crops
  int maps_for = -1;
  const void* maps_patches = nullptr;
  CUtensorMap a_patches, a_xn, a_big, a_q, a_kv;
  // CLS-only last block: strided CLS rows of Xn, compact [n][D] / [n][4D] operands
  CUtensorMap a_xn_cls, a_ocls, a_xnc, a_hc, m_kv_last, m_q_last;
  bool cls_prune = true;
  // LayerNorm folded into the consuming GEMMs (fp16 residual stream + vpf_vit_set_ln_fold)
  int ln_fold = 0;               // 0 off, 1 row statistics by rowstats16_kernel, 2 by the residual
                                 // GEMMs' epilogues (kernel only after the patch embedding)
  bool has_fold = false;
  float2* stats_final = nullptr; // [max_rows] (mean, rstd)
  float2* stats_part = nullptr;  // [parts][max_rows] (sum x, sum x^2)
  bool parts_ready = false;      // stats_part describes the current X
  CUtensorMap a_x, a_x_cls, m_kv_last_f, m_q_last_f;
  // optional in-step kernel timing (vpf_vit_set_timing): an event pair around every launch of a
  // forward, summed per kind by vpf_vit_timing
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kind;      // kind of the i-th recorded interval (events 2i, 2i+1)
  size_t ev_used = 0;
};

// kinds of vpf_vit_timing's output vector
enum { kTPatch = 0, kTLayerNorm, kTQkv, kTAttention, kTProj, kTFc1, kTFc2, kTTail, kTKinds };

struct TimedLaunch {
  vpf_vit* h; cudaStream_t s; bool on;
  TimedLaunch(vpf_vit* h_, int kind, cudaStream_t s_) : h(h_), s(s_), on(h_->timing) {
    if (!on) return;
    while (h->ev_pool.size() < h->ev_used + 2) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      h->ev_pool.push_back(e);
    }
    h->ev_kind.push_back(kind);
    cudaEventRecord(h->ev_pool[h->ev_used], s);
  }
  ~TimedLaunch() {
    if (!on) return;
    cudaEventRecord(h->ev_pool[h->ev_used + 1], s);
    h->ev_used += 2;
  }
};
#define VPF_TIMED(kind) TimedLaunch timed_launch__(h, kind, s)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// compact [n][..] operands of the CLS-only last block: q, o, xn (16-bit [n][D]), hc (16-bit
// [n][4D]), xc (f32 [n][D]), each 256-byte aligned
static size_t cls_bytes(size_t n, size_t D) {
  return 3 * align_up(n * D * 2, 256) + align_up(n * 4 * D * 2, 256) + align_up(n * D * 4, 256);
}

// crop_size 0 means the default 224; otherwise a multiple of 16 in [16, 224]
static int tokens_of(const vpf_vit_config* cfg) {
  const int s = cfg->crop_size ? cfg->crop_size : 224;
  if (s < 16 || s > 224 || s % 16) return 0;
  return (s / 16) * (s / 16) + 1;
}

extern "C" {

int vpf_abi_version(void) { return VPF_ABI_VERSION; }
const char* vpf_last_error(void) { return g_err; }

int vpf_vit_create(vpf_vit** out, const vpf_vit_config* cfg) {
  VPF_REQUIRE(out && cfg, "vpf_vit_create: null argument");
  VPF_REQUIRE(cfg->dim > 0 && cfg->dim % 64 == 0 && cfg->dim <= 1024 && cfg->depth > 0 &&
                  cfg->heads > 0 && cfg->heads * 64 == cfg->dim && cfg->max_crops > 0,
              "vpf_vit_create: need dim = 64*heads <= 1024, depth > 0, max_crops > 0");
  VPF_REQUIRE(cfg->dtype == VPF_DTYPE_BF16 || cfg->dtype == VPF_DTYPE_F16,
              "vpf_vit_create: dtype must be VPF_DTYPE_BF16 or VPF_DTYPE_F16");
  VPF_REQUIRE(tokens_of(cfg) > 0, "vpf_vit_create: crop_size %d must be 0 (= 224) or a multiple of 16 in [16, 224]",
              cfg->crop_size);
  VPF_REQUIRE(cfg->residual_dtype == VPF_RESIDUAL_F32 || cfg->residual_dtype == VPF_RESIDUAL_F16,
              "vpf_vit_create: residual_dtype must be VPF_RESIDUAL_F32 or VPF_RESIDUAL_F16");
  if (!pick_block_n(cfg->dim) || !pick_block_n(3 * cfg->dim)) {
    set_error("vpf_vit_create: dim %d unsupported (dim and 3*dim must be multiples of 128)", cfg->dim);
    return VPF_ERR_UNSUPPORTED;
  }
  vpf_vit* h = new vpf_vit();
  h->cfg = *cfg;
  h->tokens = tokens_of(cfg);
  h->ppc = h->tokens - 1;
  h->x16 = cfg->residual_dtype == VPF_RESIDUAL_F16;
  h->layers.resize(cfg->depth);
  {
    const char* e = getenv("VPF_CLS_PRUNE");    // VPF_CLS_PRUNE=0: run the last block on all rows
    h->cls_prune = !(e && e[0] == '0');
  }
  *out = h;
  return VPF_OK;
}

int vpf_vit_destroy(vpf_vit* h) {
  if (h) for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  delete h;
  return VPF_OK;
}

size_t vpf_vit_workspace_bytes(const vpf_vit_config* cfg) {
  if (!cfg || !tokens_of(cfg)) return 0;
  const size_t rows = (size_t)cfg->max_crops * tokens_of(cfg), D = cfg->dim;
  const size_t xsz = cfg->residual_dtype == VPF_RESIDUAL_F16 ? 2 : 4;
  return align_up(rows * D * xsz, 1024) + align_up(rows * D * 2, 1024) +
         align_up(rows * 4 * D * 2, 1024) + align_up(cls_bytes(cfg->max_crops, D), 1024) +
         align_up(rows * sizeof(float2) * (size_t)(1 + ln_stats_parts(cfg->dim)), 1024) + 1024;
}

int vpf_vit_set_weights(vpf_vit* h, const void* const* p, int n) {
  VPF_REQUIRE(h && p, "vpf_vit_set_weights: null argument");
  const int D = h->cfg.dim, L = h->cfg.depth;
  VPF_REQUIRE(n == VPF_VIT_N_GLOBAL + 12 * L + 2, "vpf_vit_set_weights: expected %d pointers, got %d",
              VPF_VIT_N_GLOBAL + 12 * L + 2, n);
  for (int i = 0; i < n; ++i) VPF_REQUIRE(p[i] != nullptr, "vpf_vit_set_weights: pointer %d is null", i);
  int k = 0;
  h->patch_w = p[k++]; h->patch_b = (const float*)p[k++];
  h->cls = (const float*)p[k++]; h->pos = (const float*)p[k++];
  int rc;
  if ((rc = make_map(&h->m_patch, h->patch_w, D, 768, pick_block_n(D) / gemm_ctas()))) return rc;
  for (int l = 0; l < L; ++l) {
    LayerW& w = h->layers[l];
    w.ln1_g = (const float*)p[k++]; w.ln1_b = (const float*)p[k++];
    w.qkv_w = p[k++]; w.qkv_b = (const float*)p[k++];
    w.proj_w = p[k++]; w.proj_b = (const float*)p[k++];
    w.ln2_g = (const float*)p[k++]; w.ln2_b = (const float*)p[k++];
    w.fc1_w = p[k++]; w.fc1_b = (const float*)p[k++];
    w.fc2_w = p[k++]; w.fc2_b = (const float*)p[k++];
    if ((rc = make_map(&w.m_qkv, w.qkv_w, 3 * D, D, pick_block_n(3 * D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&w.m_proj, w.proj_w, D, D, pick_block_n(D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&w.m_fc1, w.fc1_w, 4 * D, D, pick_block_n(4 * D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&w.m_fc2, w.fc2_w, D, 4 * D, pick_block_n(D) / gemm_ctas()))) return rc;
  }
  h->norm_g = (const float*)p[k++]; h->norm_b = (const float*)p[k++];
  {  // last block, CLS-only path: separate views of the fused qkv weight (Q rows, then K|V rows)
    const LayerW& w = h->layers[L - 1];
    const __nv_bfloat16* qkv = (const __nv_bfloat16*)w.qkv_w;
    if ((rc = make_map(&h->m_q_last, qkv, D, D, pick_block_n(D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&h->m_kv_last, qkv + (size_t)D * D, 2 * D, D,
                       pick_block_n(2 * D) / gemm_ctas()))) return rc;
  }
  h->has_weights = true;
  return VPF_OK;
}

// VPF_LN_FOLD: 0 = keep the LayerNorm kernels, kernel = folded GEMMs with rowstats16_kernel before
// each, anything else / unset = folded GEMMs fed by the residual epilogues' partial sums.
static int ln_fold_mode() {
  const char* e = getenv("VPF_LN_FOLD");
  if (!e) return 2;
  if (e[0] == '0') return 0;
  return e[0] == 'k' ? 1 : 2;
}

int vpf_vit_set_ln_fold(vpf_vit* h, const void* const* p, int n) {
  VPF_REQUIRE(h && p, "vpf_vit_set_ln_fold: null argument");
  const int D = h->cfg.dim, L = h->cfg.depth;
  VPF_REQUIRE(h->has_weights, "vpf_vit_set_ln_fold: call vpf_vit_set_weights first");
  VPF_REQUIRE(h->x16, "vpf_vit_set_ln_fold: needs the fp16 residual stream (VPF_RESIDUAL_F16): the "
                      "stream itself becomes the GEMM's A operand");
  VPF_REQUIRE(n == 4 * L, "vpf_vit_set_ln_fold: expected %d pointers, got %d", 4 * L, n);
  for (int i = 0; i < n; ++i) VPF_REQUIRE(p[i] != nullptr, "vpf_vit_set_ln_fold: pointer %d is null", i);
  int k = 0, rc;
  for (int l = 0; l < L; ++l) {
    LayerW& w = h->layers[l];
    w.qkv_wf = p[k++]; w.qkv_bf = (const float*)p[k++];
    w.fc1_wf = p[k++]; w.fc1_bf = (const float*)p[k++];
    if ((rc = make_map(&w.m_qkv_f, w.qkv_wf, 3 * D, D, pick_block_n(3 * D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&w.m_fc1_f, w.fc1_wf, 4 * D, D, pick_block_n(4 * D) / gemm_ctas()))) return rc;
  }
  {
    const __nv_bfloat16* qkv = (const __nv_bfloat16*)h->layers[L - 1].qkv_wf;
    if ((rc = make_map(&h->m_q_last_f, qkv, D, D, pick_block_n(D) / gemm_ctas()))) return rc;
    if ((rc = make_map(&h->m_kv_last_f, qkv + (size_t)D * D, 2 * D, D,
                       pick_block_n(2 * D) / gemm_ctas()))) return rc;
  }
  h->has_fold = true;
  h->ln_fold = ln_fold_mode();
  return VPF_OK;
}

int vpf_vit_ln_fold(const vpf_vit* h) { return h ? (h->has_fold ? h->ln_fold : 0) : 0; }

int vpf_vit_set_workspace(vpf_vit* h, void* ws, size_t bytes) {
  VPF_REQUIRE(h && ws, "vpf_vit_set_workspace: null argument");
  VPF_REQUIRE(((uintptr_t)ws & 1023) == 0, "vpf_vit_set_workspace: workspace must be 1024-byte aligned");
  VPF_REQUIRE(bytes >= vpf_vit_workspace_bytes(&h->cfg),
              "vpf_vit_set_workspace: need %zu bytes, got %zu", vpf_vit_workspace_bytes(&h->cfg), bytes);
  const size_t rows = (size_t)h->cfg.max_crops * h->tokens, D = h->cfg.dim;
  uint8_t* b = (uint8_t*)ws;
  h->X = (float*)b;            b += align_up(rows * D * (h->x16 ? 2 : 4), 1024);
  h->Xn = (__nv_bfloat16*)b;   b += align_up(rows * D * 2, 1024);
  h->big = (__nv_bfloat16*)b;  b += align_up(rows * 4 * D * 2, 1024);
  h->cbuf = b;                 b += align_up(cls_bytes(h->cfg.max_crops, D), 1024);
  h->stats_final = (float2*)b; b += rows * sizeof(float2);
  h->stats_part = (float2*)b;
  h->has_ws = true;
  h->maps_for = -1;
  return VPF_OK;
}

int vpf_vit_launches(const vpf_vit* h) {
  if (!h) return 0;
  const int L = h->cfg.depth;
  const int fold = h->has_fold ? h->ln_fold : 0;
  // per full block: 2 LayerNorms + qkv, attention, proj, fc1, fc2; folded: the LayerNorms go
  // (mode 2: one rowstats launch per forward; mode 1: one per former LayerNorm)
  const int per_block = fold == 2 ? 5 : 7;
  const int extra = fold == 2 ? 1 : 0;
  if (!h->cls_prune) return 2 /*cls, patch embed*/ + per_block * L + extra + 1 /*final LN*/;
  return 2 + per_block * (L - 1) + extra + (fold == 2 ? 9 : 10);
}

// ---- LayerNorm folded into qkv / fc1: where the row statistics come from ---------------------
// Mode 2: the residual GEMM that last rewrote X left per-column-group partial sums (parts_ready);
// otherwise (after the patch embedding, or mode 1) rowstats16_kernel makes final (mean, rstd).
static int ensure_row_stats(vpf_vit* h, int M, cudaStream_t s) {
  if (h->parts_ready) return VPF_OK;
  VPF_TIMED(kTLayerNorm);
  VPF_CUDA(launch_rowstats16(s, h->X, h->stats_final, M, h->cfg.dim));
  return VPF_OK;
}
static void set_ln_stats(const vpf_vit* h, GemmArgs& g) {
  if (h->parts_ready) {
    g.stats = h->stats_part;
    g.stats_parts = ln_stats_parts(h->cfg.dim);
    g.stats_ld = (size_t)h->cfg.max_crops * h->tokens;
  } else {
    g.stats = h->stats_final;
    g.stats_parts = 0;
  }
}
// a residual GEMM over every row of X: in mode 2 its epilogue also produces the next statistics
static void set_stats_out(vpf_vit* h, GemmArgs& g) {
  const bool epi = h->has_fold && h->ln_fold == 2;
  if (epi) {
    g.stats_out = h->stats_part;
    g.stats_out_ld = (size_t)h->cfg.max_crops * h->tokens;
  }
  h->parts_ready = epi;
}

// One full encoder block on every token row.
static int run_block(vpf_vit* h, const LayerW& w, int n, cudaStream_t s) {
  const int D = h->cfg.dim, H = h->cfg.heads, M = n * h->tokens;
  const int f16 = h->cfg.dtype == VPF_DTYPE_F16;
  const int ln_blocks = (M + 7) / 8;
  const int fold = h->has_fold ? h->ln_fold : 0;
  int rc;
  if (fold) {
    if ((rc = ensure_row_stats(h, M, s))) return rc;
    VPF_TIMED(kTQkv);
    GemmArgs g{M, 3 * D, D, w.qkv_bf, h->big, nullptr};
    g.f16 = f16;
    set_ln_stats(h, g);
    if ((rc = launch_gemm(h->a_x, w.m_qkv_f, g, kLnBiasBf16, s))) return rc;
  } else {
    {
      VPF_TIMED(kTLayerNorm);
      VPF_CUDA(launch_layernorm<__nv_bfloat16>(ln_blocks, s, h->X, D, w.ln1_g, w.ln1_b, h->Xn, M, D, f16 != 0, h->x16));
    }
    VPF_TIMED(kTQkv);
    GemmArgs g{M, 3 * D, D, w.qkv_b, h->big, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_xn, w.m_qkv, g, kBiasBf16, s))) return rc;
  }
  {
    VPF_TIMED(kTAttention);
    if ((rc = launch_attention(h->big, h->Xn, n, H, h->tokens, &h->a_q, &h->a_kv, s, f16))) return rc;
  }
  {
    VPF_TIMED(kTProj);
    GemmArgs g{M, D, D, w.proj_b, h->X, nullptr};
      g.f16 = f16;
      g.x16 = h->x16;
    set_stats_out(h, g);
    if ((rc = launch_gemm(h->a_xn, w.m_proj, g, kBiasResidualF32, s))) return rc;
  }
  if (fold) {
    if ((rc = ensure_row_stats(h, M, s))) return rc;
    VPF_TIMED(kTFc1);
    GemmArgs g{M, 4 * D, D, w.fc1_bf, h->big, nullptr};
    g.f16 = f16;
    set_ln_stats(h, g);
    if ((rc = launch_gemm(h->a_x, w.m_fc1_f, g, kLnBiasGeluBf16, s))) return rc;
  } else {
    {
      VPF_TIMED(kTLayerNorm);
      VPF_CUDA(launch_layernorm<__nv_bfloat16>(ln_blocks, s, h->X, D, w.ln2_g, w.ln2_b, h->Xn, M, D, f16 != 0, h->x16));
    }
    VPF_TIMED(kTFc1);
    GemmArgs g{M, 4 * D, D, w.fc1_b, h->big, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_xn, w.m_fc1, g, kBiasGeluBf16, s))) return rc;
  }
  {
    VPF_TIMED(kTFc2);
    GemmArgs g{M, D, 4 * D, w.fc2_b, h->X, nullptr};
      g.f16 = f16;
      g.x16 = h->x16;
    set_stats_out(h, g);
    if ((rc = launch_gemm(h->a_big, w.m_fc2, g, kBiasResidualF32, s))) return rc;
  }
  return VPF_OK;
}

// Last block, CLS rows only (exact): keys / values for every token, everything else for the n
// CLS rows.  Compact operands live in their own workspace region (`cbuf`).
struct ClsBufs {
  __nv_bfloat16 *q, *o, *xn, *hc;
  float* xc;
};
static ClsBufs cls_bufs(const vpf_vit* h, int n) {
  const size_t D = h->cfg.dim;
  uint8_t* b = h->cbuf;
  ClsBufs c;
  c.q = (__nv_bfloat16*)b;  b += align_up((size_t)n * D * 2, 256);
  c.o = (__nv_bfloat16*)b;  b += align_up((size_t)n * D * 2, 256);
  c.xn = (__nv_bfloat16*)b; b += align_up((size_t)n * D * 2, 256);
  c.hc = (__nv_bfloat16*)b; b += align_up((size_t)n * 4 * D * 2, 256);
  c.xc = (float*)b;
  return c;
}

static int run_last_block_cls(vpf_vit* h, const LayerW& w, int n, float* emb, cudaStream_t s) {
  const int D = h->cfg.dim, H = h->cfg.heads, T = h->tokens, M = n * T;
  const int f16 = h->cfg.dtype == VPF_DTYPE_F16;
  const ClsBufs c = cls_bufs(h, n);
  int rc;
  if (h->has_fold && h->ln_fold)
    if ((rc = ensure_row_stats(h, M, s))) return rc;
  VPF_TIMED(kTTail);
  if (h->has_fold && h->ln_fold) {
    {  // K, V for every token straight from the stream: rows [D, 3D) of the folded qkv weight
      GemmArgs g{M, 2 * D, D, w.qkv_bf + D, h->big, nullptr};
      g.f16 = f16;
      set_ln_stats(h, g);
      if ((rc = launch_gemm(h->a_x, h->m_kv_last_f, g, kLnBiasBf16, s))) return rc;
    }
    {  // Q for the CLS rows only (strided view of X; statistics of row crop*T)
      GemmArgs g{n, D, D, w.qkv_bf, c.q, nullptr};
      g.f16 = f16;
      set_ln_stats(h, g);
      g.stats_stride = T;
      if ((rc = launch_gemm(h->a_x_cls, h->m_q_last_f, g, kLnBiasBf16, s))) return rc;
    }
  } else {
  VPF_CUDA(launch_layernorm<__nv_bfloat16>((M + 7) / 8, s, h->X, D, w.ln1_g, w.ln1_b, h->Xn, M, D, f16 != 0, h->x16));
  {  // K, V for every token: rows [D, 3D) of the fused qkv weight
    GemmArgs g{M, 2 * D, D, w.qkv_b + D, h->big, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_xn, h->m_kv_last, g, kBiasBf16, s))) return rc;
  }
  {  // Q for the CLS rows only (strided view of Xn)
    GemmArgs g{n, D, D, w.qkv_b, c.q, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_xn_cls, h->m_q_last, g, kBiasBf16, s))) return rc;
  }
  }
  VPF_CUDA(launch_pdl(pdl_for_rows((long long)n * T), cls_attention_kernel, dim3((n * H + 7) / 8), dim3(256), 0, s, c.q, h->big, c.o,
                      n * H, H, f16 != 0, T));
  VPF_CUDA(launch_pdl(pdl_for_rows((long long)n * T), gather_cls_kernel, dim3((n * (D / 4) + 255) / 256), dim3(256), 0, s, h->X, c.xc,
                      n, D, T, h->x16));
  {
    GemmArgs g{n, D, D, w.proj_b, c.xc, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_ocls, w.m_proj, g, kBiasResidualF32, s))) return rc;
  }
  VPF_CUDA(launch_layernorm<__nv_bfloat16>((n + 7) / 8, s, c.xc, D, w.ln2_g, w.ln2_b, c.xn, n, D, f16 != 0));
  {
    GemmArgs g{n, 4 * D, D, w.fc1_b, c.hc, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_xnc, w.m_fc1, g, kBiasGeluBf16, s))) return rc;
  }
  {
    GemmArgs g{n, D, 4 * D, w.fc2_b, c.xc, nullptr};
      g.f16 = f16;
    if ((rc = launch_gemm(h->a_hc, w.m_fc2, g, kBiasResidualF32, s))) return rc;
  }
  VPF_CUDA(launch_layernorm<float>((n + 7) / 8, s, c.xc, D, h->norm_g, h->norm_b, emb, n, D));
  return VPF_OK;
}

int vpf_vit_forward(vpf_vit* h, const void* patches, int n, float* emb, void* stream) {
  VPF_REQUIRE(h && patches && emb, "vpf_vit_forward: null argument");
  if (!h->has_weights || !h->has_ws) {
    set_error("vpf_vit_forward: weights or workspace not set");
    return VPF_ERR_STATE;
  }
  VPF_REQUIRE(n > 0 && n <= h->cfg.max_crops, "vpf_vit_forward: n_crops %d outside [1, %d]", n,
              h->cfg.max_crops);
  cudaStream_t s = (cudaStream_t)stream;
  const int D = h->cfg.dim, L = h->cfg.depth, T = h->tokens;
  const int M0 = n * h->ppc, M = n * T;
  int rc;
  if (h->maps_for != n || h->maps_patches != patches) {
    if ((rc = make_map(&h->a_patches, patches, M0, 768, kBlockM))) return rc;
    if ((rc = make_map(&h->a_xn, h->Xn, M, D, kBlockM))) return rc;
    if (h->x16) {               // the fp16 stream as a GEMM A operand (LayerNorm-folded GEMMs)
      if ((rc = make_map(&h->a_x, h->X, M, D, kBlockM))) return rc;
      if ((rc = make_map(&h->a_x_cls, h->X, n, D, kBlockM, (uint64_t)T * D))) return rc;
    }
    if ((rc = make_map(&h->a_big, h->big, M, 4 * D, kBlockM))) return rc;
    if (T > 64) {             // tcgen05 attention (T <= 64: the warp-per-pair kernels need no map)
      if ((rc = make_qkv_map(&h->a_q, h->big, n, D, 128, T))) return rc;
      if ((rc = make_qkv_map(&h->a_kv, h->big, n, D, keys_padded(T), T))) return rc;
    }
    const ClsBufs c = cls_bufs(h, n);
    if ((rc = make_map(&h->a_xn_cls, h->Xn, n, D, kBlockM, (uint64_t)T * D))) return rc;
    if ((rc = make_map(&h->a_ocls, c.o, n, D, kBlockM))) return rc;
    if ((rc = make_map(&h->a_xnc, c.xn, n, D, kBlockM))) return rc;
    if ((rc = make_map(&h->a_hc, c.hc, n, 4 * D, kBlockM))) return rc;
    h->maps_for = n;
    h->maps_patches = patches;
  }

  if (h->timing) { h->ev_used = 0; h->ev_kind.clear(); }
  h->parts_ready = false;
  {
    VPF_TIMED(kTPatch);
    VPF_CUDA(launch_pdl(pdl_for_rows((long long)n * T), cls_init_kernel, dim3((n * D + 255) / 256), dim3(256), 0, s, h->X, h->cls,
                        h->pos, n, D, T, h->x16));
    GemmArgs g{M0, D, 768, h->patch_b, h->X, h->pos};
    g.f16 = h->cfg.dtype == VPF_DTYPE_F16;
    g.set_ppc(h->ppc);
    g.x16 = h->x16;
    if ((rc = launch_gemm(h->a_patches, h->m_patch, g, kPatchEmbedF32, s))) return rc;
  }
  const int full_blocks = h->cls_prune ? L - 1 : L;
  for (int l = 0; l < full_blocks; ++l)
    if ((rc = run_block(h, h->layers[l], n, s))) return rc;
  if (h->cls_prune) return run_last_block_cls(h, h->layers[L - 1], n, emb, s);
  VPF_TIMED(kTTail);
  VPF_CUDA(launch_layernorm<float>((n + 7) / 8, s, h->X, (size_t)T * D, h->norm_g, h->norm_b, emb, n, D, false, h->x16));
  return VPF_OK;
}

int vpf_vit_set_timing(vpf_vit* h, int enable) {
  VPF_REQUIRE(h, "vpf_vit_set_timing: null handle");
  h->timing = enable != 0;
  h->ev_used = 0;
  h->ev_kind.clear();
  return VPF_OK;
}

int vpf_vit_timing(vpf_vit* h, float* ms_by_kind, int n_kinds) {
  VPF_REQUIRE(h && ms_by_kind && n_kinds >= kTKinds, "vpf_vit_timing: need room for %d kinds", (int)kTKinds);
  for (int k = 0; k < n_kinds; ++k) ms_by_kind[k] = 0.f;
  if (!h->timing || !h->ev_used) return VPF_OK;
  VPF_CUDA(cudaEventSynchronize(h->ev_pool[h->ev_used - 1]));
  for (size_t i = 0; i < h->ev_kind.size(); ++i) {
    float ms = 0.f;
    VPF_CUDA(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
    ms_by_kind[h->ev_kind[i]] += ms;
  }
  return VPF_OK;
}

// ---- single-stage test hooks ----------------------------------------------------------------
int vpf_gemm_bf16(const void* A, const void* W, const float* bias, void* out, const float* pos,
                  int M, int N, int K, int mode, int patches_per_crop, int residual_dtype,
                  void* stream) {
  VPF_REQUIRE(A && W && bias && out && M > 0 && N > 0 && K > 0, "vpf_gemm_bf16: bad arguments");
  VPF_REQUIRE(K % kBlockK == 0, "vpf_gemm_bf16: K=%d must be a multiple of 64", K);
  VPF_REQUIRE(mode != kPatchEmbedF32 || (pos && patches_per_crop > 0),
              "vpf_gemm_bf16: patch-embed mode needs pos and patches_per_crop > 0");
  const int bn = pick_block_n(N);
  if (!bn) { set_error("vpf_gemm_bf16: N=%d must be a multiple of 128", N); return VPF_ERR_UNSUPPORTED; }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, A, M, K, kBlockM))) return rc;
  if ((rc = make_map(&mb, W, N, K, bn / gemm_ctas()))) return rc;
  GemmArgs g{M, N, K, bias, out, pos};
  if (mode == kPatchEmbedF32) g.set_ppc(patches_per_crop);
  g.x16 = (mode >= kBiasResidualF32 && residual_dtype == VPF_RESIDUAL_F16) ? 1 : 0;
  return launch_gemm(ma, mb, g, mode, (cudaStream_t)stream);
}

int vpf_ln_stats_parts(int D) { return ln_stats_parts(D); }

int vpf_rowstats16(const void* x16, float* stats, int rows, int D, void* stream) {
  VPF_REQUIRE(x16 && stats && rows > 0 && D > 0 && D % 8 == 0 && D <= 1024,
              "vpf_rowstats16: bad arguments (D multiple of 8, <= 1024)");
  VPF_CUDA(launch_rowstats16((cudaStream_t)stream, x16, (float2*)stats, rows, D));
  return VPF_OK;
}

int vpf_gemm_ln(const void* X16, const void* Wf, const float* bias_f, const float* stats,
                int stats_parts, void* out, int M, int N, int K, int gelu, int dtype, void* stream) {
  VPF_REQUIRE(X16 && Wf && bias_f && stats && out && M > 0 && N > 0 && K > 0,
              "vpf_gemm_ln: bad arguments");
  VPF_REQUIRE(K % kBlockK == 0, "vpf_gemm_ln: K=%d must be a multiple of 64", K);
  VPF_REQUIRE(stats_parts >= 0 && stats_parts <= 16, "vpf_gemm_ln: stats_parts outside [0, 16]");
  const int bn = pick_block_n(N);
  if (!bn) { set_error("vpf_gemm_ln: N=%d must be a multiple of 128", N); return VPF_ERR_UNSUPPORTED; }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, X16, M, K, kBlockM))) return rc;
  if ((rc = make_map(&mb, Wf, N, K, bn / gemm_ctas()))) return rc;
  GemmArgs g{M, N, K, bias_f, out, nullptr};
  g.f16 = dtype == VPF_DTYPE_F16;
  g.stats = (const float2*)stats;
  g.stats_parts = stats_parts;
  g.stats_ld = (size_t)M;
  return launch_gemm(ma, mb, g, gelu ? kLnBiasGeluBf16 : kLnBiasBf16, (cudaStream_t)stream);
}

int vpf_gemm_residual_stats(const void* A, const void* W, const float* bias, void* x16,
                            float* stats_out, int M, int N, int K, int dtype, void* stream) {
  VPF_REQUIRE(A && W && bias && x16 && stats_out && M > 0 && N > 0 && K > 0,
              "vpf_gemm_residual_stats: bad arguments");
  VPF_REQUIRE(K % kBlockK == 0, "vpf_gemm_residual_stats: K=%d must be a multiple of 64", K);
  const int bn = pick_block_n(N);
  if (!bn) { set_error("vpf_gemm_residual_stats: N=%d must be a multiple of 128", N); return VPF_ERR_UNSUPPORTED; }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, A, M, K, kBlockM))) return rc;
  if ((rc = make_map(&mb, W, N, K, bn / gemm_ctas()))) return rc;
  GemmArgs g{M, N, K, bias, x16, nullptr};
  g.f16 = dtype == VPF_DTYPE_F16;
  g.x16 = 1;
  g.stats_out = (float2*)stats_out;
  g.stats_out_ld = (size_t)M;
  return launch_gemm(ma, mb, g, kBiasResidualF32, (cudaStream_t)stream);
}

int vpf_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, int rows,
                  int D, void* stream) {
  VPF_REQUIRE(x && gamma && beta && out_bf16 && rows > 0 && D > 0 && D % 4 == 0 && D <= 1024,
              "vpf_layernorm: bad arguments (D multiple of 4, <= 1024)");
  VPF_CUDA(launch_layernorm<__nv_bfloat16>((rows + 7) / 8, (cudaStream_t)stream, x, D, gamma, beta, (__nv_bfloat16*)out_bf16, rows, D));
  return VPF_OK;
}

int vpf_attention(const void* qkv, void* out, int n_crops, int heads, int tokens, void* stream) {
  VPF_REQUIRE(qkv && out && n_crops > 0 && heads > 0 && tokens >= 2 && tokens <= kMaxTokens,
              "vpf_attention: bad arguments (2 <= tokens <= 197)");
  CUtensorMap mq, mkv;
  int rc;
  if (tokens > 64) {
    if ((rc = make_qkv_map(&mq, qkv, n_crops, heads * kHeadDim, 128, tokens))) return rc;
    if ((rc = make_qkv_map(&mkv, qkv, n_crops, heads * kHeadDim, keys_padded(tokens), tokens))) return rc;
  }
  return launch_attention((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, n_crops, heads, tokens,
                          &mq, &mkv, (cudaStream_t)stream);
}

}  // extern "C"
