// K1: per-particle crop + bilinear resize + normalise + patchify.
// Oracle: oracle/crop.py crop_resize + patchify (upstream anchor
// README.md:7, the input side of ViT feature extraction).
//
// HBM-write bound by construction: 3*S*S bf16 written per particle (301 056 B at the default crop
// size S = 224; S is a run-time multiple of 16, see SURVEY.md section 8f-2 "native-resolution
// crops"); the uint8 frame (<= 6.2 MB) lives in L2.  One CTA produces two patch rows (32 x S
// output pixels) of one crop:
//   1. the particle box and the S horizontal taps (x0, frac) -- identical for every output
//      row of the crop -- are staged in shared memory once per CTA;
//   2. the source rows the CTA touches are copied to shared memory with coalesced 4-byte loads
//      (falls back to direct global taps when the box is too large to stage);
//   3. one warp writes one 16x16 patch per iteration as three 512-byte contiguous runs (one per
//      channel), each lane producing 8 consecutive pixels (one 16-byte store per channel).
// The arithmetic mirrors the oracle's fp32 operation order with explicit round-to-nearest
// intrinsics (no FMA contraction), so the result equals bf16(oracle) bit for bit.
#include "common.cuh"

namespace vpf {

static constexpr int kMaxCrop = 224;
static constexpr int kPatch = 16;
static constexpr int kPatchK = 768;
static constexpr int kRowsPerCta = 2 * kPatch;            // 32 output rows = two patch rows
static constexpr int kStageWords = 9 * 1024;              // 36 KB of staged source rows

struct Box {
  float cx, cy, w, h;
};

__device__ __forceinline__ float sample_coord(float base, float step, int k, float hi) {
  float x = __fadd_rn(base, __fmul_rn((float)k + 0.5f, step));
  return fminf(fmaxf(x, 0.0f), hi);
}

template <bool kFromState>
__global__ void __launch_bounds__(256) crop_kernel(const uint8_t* __restrict__ frame, int H, int W,
                                                    const float* __restrict__ src, int n, int ld,
                                                    float w0, float h0,
                                                    __nv_bfloat16* __restrict__ out, bool f16,
                                                    int S /* output crop size, multiple of 16 */) {
  __shared__ uint32_t stage[kStageWords];
  __shared__ float s_fx[kMaxCrop];
  __shared__ int s_x0[kMaxCrop];
  const int grid_p = S / kPatch;                         // patches per row / column
  const int crop_patches = grid_p * grid_p;
  const int crop = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Box b;
  if (kFromState) {
    float es = expf(src[4 * (size_t)ld + crop]);
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
  const float c255 = 2.0f / 255.0f;
  const int i0 = blockIdx.y * kRowsPerCta;
  const int n_rows = min(kRowsPerCta, S - i0);           // 32, or 16 for an odd last patch row
  const int first_patch = 2 * blockIdx.y * grid_p;
  const int n_patches = (n_rows / kPatch) * grid_p;

  // ---- horizontal taps, shared by every output row ----
  if (threadIdx.x < S) {
    const float xs = sample_coord(basex, stepx, threadIdx.x, xmax);
    const float x0f = floorf(xs);
    s_fx[threadIdx.x] = __fsub_rn(xs, x0f);
    s_x0[threadIdx.x] = (int)x0f;
  }
  // ---- source window of this CTA (coordinates are monotone in the output index) ----
  const int y_lo = (int)floorf(sample_coord(basey, stepy, i0, ymax));
  const int y_hi = min((int)floorf(sample_coord(basey, stepy, i0 + n_rows - 1, ymax)) + 1, H - 1);
  const int x_lo = (int)floorf(sample_coord(basex, stepx, 0, xmax));
  const int x_hi = min((int)floorf(sample_coord(basex, stepx, S - 1, xmax)) + 1, W - 1);
  const int rows = y_hi - y_lo + 1;
  const int row_bytes = (x_hi - x_lo + 1) * 3;
  const int pitch = (row_bytes + 6) / 4 + 1;                  // words per staged row
  const bool staged = rows * pitch <= kStageWords;
  if (staged) {
    const uintptr_t end = reinterpret_cast<uintptr_t>(frame) + (size_t)H * W * 3;
    for (int idx = threadIdx.x; idx < rows * pitch; idx += 256) {
      const int r = idx / pitch, wd = idx - r * pitch;
      const uintptr_t g0 = reinterpret_cast<uintptr_t>(frame) + ((size_t)(y_lo + r) * W + x_lo) * 3;
      const uintptr_t ga = (g0 & ~(uintptr_t)3) + (uintptr_t)wd * 4;   // 4-byte aligned word
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
  const uint8_t* sbytes = reinterpret_cast<const uint8_t*>(stage);
  const int ky = lane >> 1, kx0 = (lane & 1) * 8;

  for (int pl = warp; pl < n_patches; pl += 8) {
    const int p = first_patch + pl;
    const int py = p / grid_p, px = p - py * grid_p;
    const int i = py * kPatch + ky;
    const float ys = sample_coord(basey, stepy, i, ymax);
    const float y0f = floorf(ys);
    const float fy = __fsub_rn(ys, y0f);
    const float gy = __fsub_rn(1.0f, fy);
    const int y0 = (int)y0f, y1 = min(y0 + 1, H - 1);
    const uint8_t* tap = staged ? sbytes : frame;
    long long off0, off1;
    if (staged) {
      // a staged row starts at the 4-byte aligned address at or below its first pixel (x_lo)
      const int sh0 = (int)((reinterpret_cast<uintptr_t>(frame) + ((size_t)y0 * W + x_lo) * 3) & 3);
      const int sh1 = (int)((reinterpret_cast<uintptr_t>(frame) + ((size_t)y1 * W + x_lo) * 3) & 3);
      off0 = (long long)(y0 - y_lo) * pitch * 4 + sh0 - x_lo * 3;
      off1 = (long long)(y1 - y_lo) * pitch * 4 + sh1 - x_lo * 3;
    } else {
      off0 = (long long)y0 * W * 3;
      off1 = (long long)y1 * W * 3;
    }
    __align__(16) uint16_t v[3][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = px * kPatch + kx0 + e;
      const float fx = s_fx[j];
      const int x0 = s_x0[j], x1 = min(x0 + 1, W - 1);
      const float gx = __fsub_rn(1.0f, fx);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float p00 = (float)tap[off0 + x0 * 3 + c], p01 = (float)tap[off0 + x1 * 3 + c];
        float p10 = (float)tap[off1 + x0 * 3 + c], p11 = (float)tap[off1 + x1 * 3 + c];
        float top = __fadd_rn(__fmul_rn(gx, p00), __fmul_rn(fx, p01));
        float bot = __fadd_rn(__fmul_rn(gx, p10), __fmul_rn(fx, p11));
        float val = __fadd_rn(__fmul_rn(gy, top), __fmul_rn(fy, bot));
        v[c][e] = pack16(__fsub_rn(__fmul_rn(val, c255), 1.0f), f16);
      }
    }
    __nv_bfloat16* dst = out + ((size_t)crop * crop_patches + p) * kPatchK + ky * kPatch + kx0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<uint4*>(dst + c * 256) = *reinterpret_cast<const uint4*>(v[c]);
  }
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
  crop_kernel<true><<<dim3(n, (crop_size + kRowsPerCta - 1) / kRowsPerCta), 256, 0,
                      (cudaStream_t)stream>>>(frame, H, W, state, n, ld, w0, h0,
                                              (__nv_bfloat16*)out_bf16, dtype == VPF_DTYPE_F16,
                                              crop_size);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_crop_boxes(const uint8_t* frame, int H, int W, const float* boxes, int n, void* out_bf16,
                   int dtype, int crop_size, void* stream) {
  VPF_REQUIRE(dtype == VPF_DTYPE_BF16 || dtype == VPF_DTYPE_F16, "vpf_crop_boxes: bad dtype");
  VPF_REQUIRE(crop_size_ok(crop_size), "vpf_crop_boxes: crop_size %d must be a multiple of 16 in [16, 224]", crop_size);
  VPF_REQUIRE(frame && boxes && out_bf16 && H > 0 && W > 0 && n > 0,
              "vpf_crop_boxes: bad arguments");
  VPF_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "vpf_crop_boxes: out must be 16-byte aligned");
  crop_kernel<false><<<dim3(n, (crop_size + kRowsPerCta - 1) / kRowsPerCta), 256, 0,
                       (cudaStream_t)stream>>>(frame, H, W, boxes, n, 0, 1.f, 1.f,
                                               (__nv_bfloat16*)out_bf16, dtype == VPF_DTYPE_F16,
                                               crop_size);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

}  // extern "C"
