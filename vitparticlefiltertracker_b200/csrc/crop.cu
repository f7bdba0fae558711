// K1: per-particle crop + bilinear resize + normalise + patchify.
// Oracle: oracle/crop.py crop_resize + patchify (upstream anchor
// README.md:7, the input side of ViT feature extraction).
//
// HBM-write bound: 3*224*224 bf16 = 301 056 B written per particle; the uint8 frame (<= 6.2 MB)
// is read through L1/L2.  One CTA handles PATCHES_PER_CTA patches of one crop; one warp writes
// one 16x16 patch per iteration as three 512-byte contiguous runs (one per channel), each lane
// producing 8 consecutive pixels (one 16-byte store per channel).  The arithmetic mirrors the
// oracle's fp32 operation order with explicit round-to-nearest intrinsics (no FMA contraction),
// so the result equals bf16(oracle) bit for bit.
#include "common.cuh"

namespace vpf {

static constexpr int kCrop = 224;
static constexpr int kPatch = 16;
static constexpr int kGridP = 14;
static constexpr int kPatches = 196;
static constexpr int kPatchK = 768;
static constexpr int kSplit = 7;                       // CTAs per crop
static constexpr int kPatchesPerCta = kPatches / kSplit;  // 28 = two patch rows

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
                                                    __nv_bfloat16* __restrict__ out) {
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
  const float stepx = __fdiv_rn(b.w, (float)kCrop), stepy = __fdiv_rn(b.h, (float)kCrop);
  const float basex = __fsub_rn(__fsub_rn(b.cx, __fmul_rn(0.5f, b.w)), 0.5f);
  const float basey = __fsub_rn(__fsub_rn(b.cy, __fmul_rn(0.5f, b.h)), 0.5f);
  const float xmax = (float)(W - 1), ymax = (float)(H - 1);
  const int ky = lane >> 1, kx0 = (lane & 1) * 8;
  const float c255 = 2.0f / 255.0f;

  for (int pl = warp; pl < kPatchesPerCta; pl += 8) {
    const int p = blockIdx.y * kPatchesPerCta + pl;
    const int py = p / kGridP, px = p - py * kGridP;
    const int i = py * kPatch + ky;
    const float ys = sample_coord(basey, stepy, i, ymax);
    const float y0f = floorf(ys);
    const float fy = __fsub_rn(ys, y0f);
    const int y0 = (int)y0f, y1 = min(y0 + 1, H - 1);
    const uint8_t* row0 = frame + (size_t)y0 * W * 3;
    const uint8_t* row1 = frame + (size_t)y1 * W * 3;
    __align__(16) __nv_bfloat16 v[3][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = px * kPatch + kx0 + e;
      const float xs = sample_coord(basex, stepx, j, xmax);
      const float x0f = floorf(xs);
      const float fx = __fsub_rn(xs, x0f);
      const int x0 = (int)x0f, x1 = min(x0 + 1, W - 1);
      const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float p00 = (float)__ldg(row0 + x0 * 3 + c), p01 = (float)__ldg(row0 + x1 * 3 + c);
        float p10 = (float)__ldg(row1 + x0 * 3 + c), p11 = (float)__ldg(row1 + x1 * 3 + c);
        float top = __fadd_rn(__fmul_rn(gx, p00), __fmul_rn(fx, p01));
        float bot = __fadd_rn(__fmul_rn(gx, p10), __fmul_rn(fx, p11));
        float val = __fadd_rn(__fmul_rn(gy, top), __fmul_rn(fy, bot));
        v[c][e] = __float2bfloat16_rn(__fsub_rn(__fmul_rn(val, c255), 1.0f));
      }
    }
    __nv_bfloat16* dst = out + ((size_t)crop * kPatches + p) * kPatchK + ky * kPatch + kx0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<uint4*>(dst + c * 256) = *reinterpret_cast<const uint4*>(v[c]);
  }
}

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_crop_state(const uint8_t* frame, int H, int W, const float* state, int n, int ld,
                   float w0, float h0, void* out_bf16, void* stream) {
  VPF_REQUIRE(frame && state && out_bf16 && H > 0 && W > 0 && n > 0 && ld >= n,
              "vpf_crop_state: bad arguments");
  VPF_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "vpf_crop_state: out must be 16-byte aligned");
  crop_kernel<true><<<dim3(n, kSplit), 256, 0, (cudaStream_t)stream>>>(
      frame, H, W, state, n, ld, w0, h0, (__nv_bfloat16*)out_bf16);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_crop_boxes(const uint8_t* frame, int H, int W, const float* boxes, int n, void* out_bf16,
                   void* stream) {
  VPF_REQUIRE(frame && boxes && out_bf16 && H > 0 && W > 0 && n > 0,
              "vpf_crop_boxes: bad arguments");
  VPF_REQUIRE(((uintptr_t)out_bf16 & 15) == 0, "vpf_crop_boxes: out must be 16-byte aligned");
  crop_kernel<false><<<dim3(n, kSplit), 256, 0, (cudaStream_t)stream>>>(
      frame, H, W, boxes, n, 0, 1.f, 1.f, (__nv_bfloat16*)out_bf16);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

}  // extern "C"
