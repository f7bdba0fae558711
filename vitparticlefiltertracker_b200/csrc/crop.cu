// K1: per-particle crop + bilinear resize + normalise + patchify.
// Oracle: oracle/crop.py crop_resize + patchify (upstream anchor
// README.md:7, the input side of ViT feature extraction).
//
// HBM-write bound by construction: 3*S*S bf16 written per particle (301 056 B at the default crop
// size S = 224; S is a run-time multiple of 16, see SURVEY.md section 8f-2 "native-resolution
// crops"); the uint8 frame (<= 6.2 MB) lives in L2.  The output is written directly in the
// patch-embedding GEMM's A layout [crop][patch][c*256 + ky*16 + kx], so no im2col pass exists.
// The arithmetic keeps the oracle's fp32 operation order with explicit round-to-nearest
// operations (no FMA contraction), so the result equals bf16(oracle) bit for bit.
#include "common.cuh"

namespace vpf {

static constexpr int kMaxCrop = 224;
static constexpr int kPatch = 16;
static constexpr int kPatchK = 768;

struct Box {
  float cx, cy, w, h;
};

__device__ __forceinline__ float sample_coord(float base, float step, int k, float hi) {
  float x = __fadd_rn(base, __fmul_rn((float)k + 0.5f, step));
  return fminf(fmaxf(x, 0.0f), hi);
}

// ---------------------------------------------------------------------------------------------
// Separable bilinear.  One CTA = one patch row (16 x S output pixels) of one crop, 7 warps.
// Bilinear interpolation as the oracle writes it is  val = gy*top + fy*bot  with
//   top = gx*p[y0][x0] + fx*p[y0][x1],  bot = gx*p[y1][x0] + fx*p[y1][x1]:
// `top` / `bot` depend only on (source row, output column), and with up-sampling (box smaller than
// the crop, the tracker's normal case) the 16 output rows of a CTA touch only ~16*h/S + 2 source
// rows.  So:
//   phase A  the particle box and the S horizontal taps (x0, frac) -> shared memory
//   phase B  H[r][c][j] = gx*p[r][x0_j][c] + fx*p[r][x1_j][c] for every touched source row r, once,
//            in fp32 with the oracle's operation order -> shared memory (<= h_cap rows)
//   phase C  out[i][j][c] = gy_i*H[y0_i] + fy_i*H[y1_i], normalise, convert: packed fp32x2
//            products, never-fused scalar sums, 16-byte shared loads, one warp per 16x16 patch
//            writing three 512-byte runs.
// Same roundings in the same order as the oracle => still bit-equal to bf16(oracle); about 8
// instructions per output value instead of the 20 of a per-output four-tap kernel.  Patch rows that
// touch more than h_cap source rows (down-sampling: no row reuse to exploit) stage the raw source
// window as bytes instead (coalesced 4-byte loads) and interpolate per output, same arithmetic;
// windows too large even for that take their taps from global memory.
static constexpr int kV2Threads = 224;                       // 7 warps: 14 patches in two rounds
static constexpr int kHRowsMax = 20;                         // largest H-row budget a launch may ask for
static constexpr int kHPitchMax = 3 * kMaxCrop + 4;          // floats per H row (+4: bank spread)
static constexpr int kV2SmemMax = kHRowsMax * kHPitchMax * 4;   // 54 080 B (dynamic)

__device__ __forceinline__ float2 fmul2_rn(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
// Two outputs: products packed (FMUL2), sums with the scalar never-fused intrinsic.  (ptxas
// contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 -- checked in the SASS -- which would change
// the oracle's roundings; __fadd_rn is documented never to be merged into an FMA.)
__device__ __forceinline__ uint32_t blend2(float2 gy2, float2 fy2, float2 top, float2 bot, bool f16) {
  const float2 t0 = fmul2_rn(gy2, top), t1 = fmul2_rn(fy2, bot);
  const float2 val = make_float2(__fadd_rn(t0.x, t1.x), __fadd_rn(t0.y, t1.y));
  const float2 nn = fmul2_rn(val, make_float2(2.0f / 255.0f, 2.0f / 255.0f));
  return pack16x2(__fadd_rn(nn.x, -1.0f), __fadd_rn(nn.y, -1.0f), f16);
}

template <bool kFromState>
__global__ void __launch_bounds__(kV2Threads) crop_sep_kernel(const uint8_t* __restrict__ frame,
                                                              int H, int W,
                                                              const float* __restrict__ src, int n,
                                                              int ld, float w0, float h0,
                                                              __nv_bfloat16* __restrict__ out,
                                                              bool f16, int S, int h_cap, int smem_words) {
  extern __shared__ __align__(16) float s_h[];               // [h_cap][3][S] (+4 floats per row)
  __shared__ float s_fx[kMaxCrop];
  __shared__ int s_x0[kMaxCrop];
  pdl_launch_dependents();
  const int grid_p = S / kPatch;
  const int crop = blockIdx.x, py = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();                                                // the particle state is the motion kernel's output
  Box b;
  if (kFromState) {
    const float es = (float)exp((double)src[4 * (size_t)ld + crop]);   // D14: fp64, rounded once
    b.cx = src[crop];
    b.cy = src[(size_t)ld + crop];
    b.w = __fmul_rn(w0, es);
    b.h = __fmul_rn(h0, es);
  } else {
    b.cx = src[crop * 4 + 0]; b.cy = src[crop * 4 + 1];
    b.w = src[crop * 4 + 2];  b.h = src[crop * 4 + 3];
  }
  const float stepx = __fdiv_rn(b.w, (float)S), stepy = __fdiv_rn(b.h, (float)S);
  const float basex = __fsub_rn(__fsub_rn(b.cx, __fmul_rn(0.5f, b.w)), 0.5f);
  const float basey = __fsub_rn(__fsub_rn(b.cy, __fmul_rn(0.5f, b.h)), 0.5f);
  const float xmax = (float)(W - 1), ymax = (float)(H - 1);
  const int i0 = py * kPatch;

  // ---- phase A: horizontal taps ----
  if (threadIdx.x < S) {
    const float xs = sample_coord(basex, stepx, threadIdx.x, xmax);
    const float x0f = floorf(xs);
    s_fx[threadIdx.x] = __fsub_rn(xs, x0f);
    s_x0[threadIdx.x] = (int)x0f;
  }
  // source rows this patch row touches (sample coordinates are monotone in the output index)
  const int y_lo = (int)floorf(sample_coord(basey, stepy, i0, ymax));
  const int y_hi = min((int)floorf(sample_coord(basey, stepy, i0 + kPatch - 1, ymax)) + 1, H - 1);
  const int rows = y_hi - y_lo + 1;
  const int pitch = 3 * S + 4;
  // 0: separable (H rows fit);  1: raw source window staged as bytes (down-sampling: no row is
  // reused, so interpolate per output from shared memory);  2: window too large, global taps
  int mode = rows <= h_cap ? 0 : 2;
  int x_lo = 0, wpitch = 0;
  if (mode == 2) {
    x_lo = (int)floorf(sample_coord(basex, stepx, 0, xmax));
    const int x_hi = min((int)floorf(sample_coord(basex, stepx, S - 1, xmax)) + 1, W - 1);
    wpitch = ((x_hi - x_lo + 1) * 3 + 6) / 4 + 1;              // words per staged row
    if (rows * wpitch <= smem_words) mode = 1;
  }
  __syncthreads();

  if (mode == 0) {
    // ---- phase B: horizontally interpolated source rows ----
    for (int idx = threadIdx.x; idx < rows * S; idx += blockDim.x) {
      const int r = idx / S, j = idx - r * S;
      const float fx = s_fx[j];
      const float gx = __fsub_rn(1.0f, fx);
      const int x0 = s_x0[j], x1 = min(x0 + 1, W - 1);
      const uint8_t* row = frame + (size_t)(y_lo + r) * W * 3;
      float* h = s_h + r * pitch + j;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p0 = (float)__ldg(row + x0 * 3 + c), p1 = (float)__ldg(row + x1 * 3 + c);
        h[c * S] = __fadd_rn(__fmul_rn(gx, p0), __fmul_rn(fx, p1));
      }
    }
  } else if (mode == 1) {
    // ---- raw window, 4-byte aligned words (coalesced) ----
    uint32_t* stage = reinterpret_cast<uint32_t*>(s_h);
    const uintptr_t end = reinterpret_cast<uintptr_t>(frame) + (size_t)H * W * 3;
    for (int idx = threadIdx.x; idx < rows * wpitch; idx += blockDim.x) {
      const int r = idx / wpitch, wd = idx - r * wpitch;
      const uintptr_t g0 = reinterpret_cast<uintptr_t>(frame) + ((size_t)(y_lo + r) * W + x_lo) * 3;
      const uintptr_t ga = (g0 & ~(uintptr_t)3) + (uintptr_t)wd * 4;
      uint32_t v = 0;
      if (ga + 4 <= end) {
        v = __ldg(reinterpret_cast<const uint32_t*>(ga));
      } else {
        for (int k = 0; k < 4; ++k)
          if (ga + k < end) v |= (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(ga + k)) << (8 * k);
      }
      stage[idx] = v;
    }
  }
  __syncthreads();

  // ---- phase C: vertical blend, normalise, convert, store (one warp per patch) ----
  const int ky = lane >> 1, kx0 = (lane & 1) * 8;
  const int i = i0 + ky;
  const float ys = sample_coord(basey, stepy, i, ymax);
  const float y0f = floorf(ys);
  const float fy = __fsub_rn(ys, y0f);
  const float gy = __fsub_rn(1.0f, fy);
  const int y0 = (int)y0f, y1 = min(y0 + 1, H - 1);
  const float2 gy2 = make_float2(gy, gy), fy2 = make_float2(fy, fy);
  for (int px = warp; px < grid_p; px += (int)(blockDim.x >> 5)) {
    const int j0 = px * kPatch + kx0;
    __nv_bfloat16* dst = out + ((size_t)crop * grid_p * grid_p + py * grid_p + px) * kPatchK +
                         ky * kPatch + kx0;
    if (mode == 0) {
      const float* h0p = s_h + (y0 - y_lo) * pitch + j0;
      const float* h1p = s_h + (y1 - y_lo) * pitch + j0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(h0p + c * S + q * 4);
          const float4 u = *reinterpret_cast<const float4*>(h1p + c * S + q * 4);
          ow[q * 2 + 0] = blend2(gy2, fy2, make_float2(t.x, t.y), make_float2(u.x, u.y), f16);
          ow[q * 2 + 1] = blend2(gy2, fy2, make_float2(t.z, t.w), make_float2(u.z, u.w), f16);
        }
        *reinterpret_cast<uint4*>(dst + c * 256) = o;
      }
    } else {
      // four taps per output value (from the staged window or from global), identical arithmetic
      const uint8_t* tap = frame;
      long long off0 = (long long)y0 * W * 3, off1 = (long long)y1 * W * 3;
      if (mode == 1) {
        // a staged row starts at the 4-byte aligned address at or below its first pixel (x_lo)
        tap = reinterpret_cast<const uint8_t*>(s_h);
        const int sh0 = (int)((reinterpret_cast<uintptr_t>(frame) + ((size_t)y0 * W + x_lo) * 3) & 3);
        const int sh1 = (int)((reinterpret_cast<uintptr_t>(frame) + ((size_t)y1 * W + x_lo) * 3) & 3);
        off0 = (long long)(y0 - y_lo) * wpitch * 4 + sh0 - x_lo * 3;
        off1 = (long long)(y1 - y_lo) * wpitch * 4 + sh1 - x_lo * 3;
      }
      __align__(16) uint16_t v[3][8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float fx = s_fx[j0 + e];
        const int x0 = s_x0[j0 + e], x1 = min(x0 + 1, W - 1);
        const float gx = __fsub_rn(1.0f, fx);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float p00 = (float)tap[off0 + x0 * 3 + c], p01 = (float)tap[off0 + x1 * 3 + c];
          const float p10 = (float)tap[off1 + x0 * 3 + c], p11 = (float)tap[off1 + x1 * 3 + c];
          const float top = __fadd_rn(__fmul_rn(gx, p00), __fmul_rn(fx, p01));
          const float bot = __fadd_rn(__fmul_rn(gx, p10), __fmul_rn(fx, p11));
          const float val = __fadd_rn(__fmul_rn(gy, top), __fmul_rn(fy, bot));
          v[c][e] = pack16(__fsub_rn(__fmul_rn(val, 2.0f / 255.0f), 1.0f), f16);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<uint4*>(dst + c * 256) = *reinterpret_cast<const uint4*>(v[c]);
    }
  }
}

template <bool kFromState>
static int launch_crop(const uint8_t* frame, int H, int W, const float* src, int n, int ld, float w0,
                       float h0, void* out, bool f16, int S, cudaStream_t stream) {
  {
    static bool attr[32] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 31]) {
      VPF_CUDA(cudaFuncSetAttribute(crop_sep_kernel<kFromState>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kV2SmemMax));
      attr[dev & 31] = true;
    }
    // H-row budget per CTA: what a box of the nominal height needs (16*h/S + 2 source rows per
    // patch row) with 40 % head-room for the scale state, so that small boxes leave shared memory
    // for more resident CTAs; explicit boxes (height unknown on the host) get the maximum.
    // Down-sampling boxes (no H-row reuse) get room for the raw byte window instead.
    int h_cap = kHRowsMax;
    int smem = h_cap * (3 * S + 4) * 4;
    if (kFromState) {
      const float rows_f = 16.0f * h0 * 1.4f / (float)S + 3.0f;
      h_cap = rows_f < 6.f ? 6 : (rows_f > (float)kHRowsMax ? kHRowsMax : (int)rows_f);
      smem = h_cap * (3 * S + 4) * 4;
      if (rows_f > (float)kHRowsMax) {
        const float raw = rows_f * (w0 * 1.4f * 3.0f + 16.0f);
        smem = raw > (float)kV2SmemMax ? kV2SmemMax : (raw > (float)smem ? (int)raw : smem);
      }
    } else if (S < kMaxCrop) {
      smem = kV2SmemMax;           // explicit boxes: box size unknown on the host
    }
    smem = (smem + 15) & ~15;
    // one warp per two patches (S = 224 -> 7 warps); when the raw window is staged instead, every
    // thread helps with the copy
    int warps = (S / kPatch + 1) / 2 < 2 ? 2 : (S / kPatch + 1) / 2;
    if (smem > h_cap * (3 * S + 4) * 4 || !kFromState) warps = kV2Threads / 32;
    VPF_CUDA(launch_pdl(pdl_for_rows((long long)n * (S / kPatch) * (S / kPatch)),
                        crop_sep_kernel<kFromState>, dim3(n, S / kPatch), dim3(warps * 32),
                        (size_t)smem, stream, frame, H, W, src, n, ld, w0, h0,
                        (__nv_bfloat16*)out, f16, S, h_cap, smem / 4));
  }
  return VPF_OK;
}

}  // namespace vpf

using namespace vpf;

extern "C" {

static bool crop_size_ok(int s) { return s >= kPatch && s <= kMaxCrop && s % kPatch == 0; }

int vpf_crop_state(const uint8_t* frame, int H, int W, const float* state, int n, int ld,
                   float w0, float h0, void* out_bf16, int dtype, int crop_size, void* stream) {
  VPF_REQUIRE(dtype == VPF_DTYPE_BF16 || dtype == VPF_DTYPE_F16, "vpf_crop_state: bad dtype");
  VPF_REQUIRE(crop_size_ok(crop_size), "vpf_crop_state: crop_size %d must be a multiple of 16 in [16, 224]", crop_size);
  VPF_REQUIRE(frame && state && out_bf16 && H > 0 && W > 0 && n > 0 && ld >= n,
              "vpf_crop_state: bad arguments");
  VPF_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "vpf_crop_state: out must be 16-byte aligned");
  return launch_crop<true>(frame, H, W, state, n, ld, w0, h0, out_bf16, dtype == VPF_DTYPE_F16,
                           crop_size, (cudaStream_t)stream);
}

int vpf_crop_boxes(const uint8_t* frame, int H, int W, const float* boxes, int n, void* out_bf16,
                   int dtype, int crop_size, void* stream) {
  VPF_REQUIRE(dtype == VPF_DTYPE_BF16 || dtype == VPF_DTYPE_F16, "vpf_crop_boxes: bad dtype");
  VPF_REQUIRE(crop_size_ok(crop_size), "vpf_crop_boxes: crop_size %d must be a multiple of 16 in [16, 224]", crop_size);
  VPF_REQUIRE(frame && boxes && out_bf16 && H > 0 && W > 0 && n > 0,
              "vpf_crop_boxes: bad arguments");
  VPF_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "vpf_crop_boxes: out must be 16-byte aligned");
  return launch_crop<false>(frame, H, W, boxes, n, 0, 1.f, 1.f, out_bf16, dtype == VPF_DTYPE_F16,
                            crop_size, (cudaStream_t)stream);
}

}  // extern "C"
