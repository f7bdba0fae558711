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
// Column-owner kernel.  One CTA = one crop; one thread = two adjacent output columns, walked down
// all S output rows.  Bilinear interpolation as the oracle writes it is
//   val = gy*top + fy*bot,   top = gx*p[y0][x0] + fx*p[y0][x1],   bot = gx*p[y1][x0] + fx*p[y1][x1].
// `top` / `bot` depend only on (source row, output column), and y0 is nondecreasing in the output
// row, so a thread keeps the two horizontally interpolated source rows of its columns in
// registers and refreshes one only when y0 advances (every S/h rows: ~2.3 output rows per source
// row for a 96-pixel box).  Everything that is per crop or per row is done ONCE per crop (the round-1
// kernel, one CTA per patch row, paid box / taps / barriers 14 times per crop and ran at 24
// instructions per output value, issue-bound at 75 % issue-active -- profiles/ncu_crop_attn_r02.json):
//   prologue   box -> per-row (y0, fy) table in shared memory; the thread's own (x0, fx) taps in
//              registers; the crop's source window (~100 x 100 pixels for a 96-pixel target)
//              staged in shared memory with coalesced aligned 4-byte loads -- in one piece when
//              it fits the launch's budget, else in bands of a few patch rows
//   main loop  per row: 2 table loads, <= 1 row refresh (12 shared byte loads, bytes -> floats with
//              the 2^23 magic constant instead of I2F, 3 x (2 FMUL2 + FADD2)), 3 x (3 FMUL2 +
//              2 FADD2 + 1 pack) for six outputs, three 4-byte stores that fill whole 32-byte
//              sectors of the patch-major output (8 lanes = one 16-pixel patch row = 32 B)
// All arithmetic keeps the oracle's fp32 operation order with round-to-nearest operations (packed
// products, scalar never-fused sums: see sum_rn), so the output still equals bf16(oracle) bit for
// bit.  Windows that do not fit the launch's shared-memory budget (boxes much
// larger than the nominal target, down-sampling) read their taps from global memory instead: same
// arithmetic, same result, slower.
static constexpr int kColThreads = 128;                      // 2 columns per thread, S/2 <= 112 active
static constexpr int kWinWordsMax = 24 * 1024;               // 96 KB: largest window a launch may ask for

__device__ __forceinline__ float2 fmul2_rn(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2_rn(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
// two bytes (0..255) -> two floats, exactly: 0x4B000000 | b is the float 2^23 + b
__device__ __forceinline__ float2 bytes2f(uint32_t a, uint32_t b) {
  return fadd2_rn(make_float2(__uint_as_float(0x4B000000u | a), __uint_as_float(0x4B000000u | b)),
                  make_float2(-8388608.0f, -8388608.0f));
}
// Sum of two packed products, each rounded on its own: products packed (FMUL2), sums with the
// scalar never-fused intrinsic.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 -- checked
// in the SASS, with and without -fmad=false -- which would skip the oracle's rounding of the
// product; __fadd_rn is documented never to be merged into an FMA.
__device__ __forceinline__ float2 sum_rn(float2 a, float2 b) {
  return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
// two outputs of one channel: gy*top + fy*bot, then (v * 2/255) - 1, packed to 16 bits
__device__ __forceinline__ uint32_t blend2(float2 gy2, float2 fy2, float2 top, float2 bot, bool f16) {
  const float2 val = sum_rn(fmul2_rn(gy2, top), fmul2_rn(fy2, bot));
  const float2 nn = fmul2_rn(val, make_float2(2.0f / 255.0f, 2.0f / 255.0f));
  return pack16x2(__fadd_rn(nn.x, -1.0f), __fadd_rn(nn.y, -1.0f), f16);
}

// Horizontally interpolated source row for the thread's two columns, three channels:
//   h[c] = gx * p[x0][c] + fx * p[x1][c]     (.x = column A, .y = column B)
// `row` points at pixel 0 of the source row (inside the shared window or in global memory).
template <bool kShared>
__device__ __forceinline__ void load_h(const uint8_t* row, int oa0, int oa1, int ob0, int ob1,
                                       float2 gx2, float2 fx2, float2 (&h)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t a0, a1, b0, b1;
    if (kShared) {
      a0 = row[oa0 + c]; a1 = row[oa1 + c]; b0 = row[ob0 + c]; b1 = row[ob1 + c];
    } else {
      a0 = __ldg(row + oa0 + c); a1 = __ldg(row + oa1 + c);
      b0 = __ldg(row + ob0 + c); b1 = __ldg(row + ob1 + c);
    }
    h[c] = sum_rn(fmul2_rn(gx2, bytes2f(a0, b0)), fmul2_rn(fx2, bytes2f(a1, b1)));
  }
}

template <bool kFromState>
__global__ void __launch_bounds__(kColThreads) crop_col_kernel(const uint8_t* __restrict__ frame,
                                                               int H, int W,
                                                               const float* __restrict__ src, int n,
                                                               int ld, float w0, float h0,
                                                               __nv_bfloat16* __restrict__ out,
                                                               bool f16, int S, int win_words) {
  extern __shared__ __align__(16) uint8_t s_dyn[];           // [S] y0 | [S] fy | window words
  int* s_y0 = reinterpret_cast<int*>(s_dyn);
  float* s_fy = reinterpret_cast<float*>(s_dyn) + S;
  uint32_t* s_win = reinterpret_cast<uint32_t*>(s_dyn) + 2 * S;
  pdl_launch_dependents();
  const int crop = blockIdx.x, tid = threadIdx.x;
  const int grid_p = S / kPatch;
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

  // ---- per-row table: source row and vertical fraction of every output row ----
  for (int i = tid; i < S; i += (int)blockDim.x) {
    const float ys = sample_coord(basey, stepy, i, ymax);
    const float y0f = floorf(ys);
    s_y0[i] = (int)y0f;
    s_fy[i] = __fsub_rn(ys, y0f);
  }
  // ---- horizontal extent of the crop's source window (sample coordinates are monotone in the
  // output index when the box size is positive; anything else reads global memory) ----
  const int x_lo = (int)floorf(sample_coord(basex, stepx, 0, xmax));
  const int x_hi = min((int)floorf(sample_coord(basex, stepx, S - 1, xmax)) + 1, W - 1);
  const int wpitch = ((x_hi - x_lo + 1) * 3 + 6) / 4 + 1;    // 4-byte words per staged row
  const bool mono = stepx >= 0.0f && stepy >= 0.0f && x_hi >= x_lo;
  const uintptr_t f0 = reinterpret_cast<uintptr_t>(frame);

  // ---- the thread's two columns: horizontal taps in registers ----
  const bool active = 2 * tid < S;
  const int ja = active ? 2 * tid : 0;
  const float xsa = sample_coord(basex, stepx, ja, xmax), xsb = sample_coord(basex, stepx, ja + 1, xmax);
  const float xa0f = floorf(xsa), xb0f = floorf(xsb);
  const float2 fx2 = make_float2(__fsub_rn(xsa, xa0f), __fsub_rn(xsb, xb0f));
  const float2 gx2 = make_float2(__fsub_rn(1.0f, fx2.x), __fsub_rn(1.0f, fx2.y));
  const int xa0 = (int)xa0f, xb0 = (int)xb0f;
  const int oa0 = xa0 * 3, oa1 = min(xa0 + 1, W - 1) * 3, ob0 = xb0 * 3, ob1 = min(xb0 + 1, W - 1) * 3;
  const int px = ja >> 4, kx = ja & 15;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) +
                 (((size_t)crop * grid_p * grid_p + px) * kPatchK + kx) * 2;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(s_win);

  float2 ht[3], hb[3];                                       // interpolated rows top_y / bot_y
  int top_y = -1, bot_y = -1;
  __syncthreads();                                           // the row table is complete

  // ---- bands: as many patch rows as the window budget holds source rows for (the whole crop for
  // a box of the nominal size; a few patch rows for larger boxes, so shared memory per CTA -- and
  // with it the number of resident CTAs -- does not depend on the box) ----
  int py = 0;
#pragma unroll 1
  while (py < grid_p) {
    // grow the band while its source rows fit
    const int y_lo = s_y0[py * kPatch];
    int band = 0, y_hi = y_lo;
    if (mono) {
      while (py + band < grid_p) {
        const int yh = min(s_y0[(py + band) * kPatch + kPatch - 1] + 1, H - 1);
        if ((long long)(yh - y_lo + 1) * wpitch > (long long)win_words) break;
        y_hi = yh;
        ++band;
      }
    }
    const bool use_win = band > 0;
    if (!use_win) band = 1;                                  // this patch row reads global memory
    if (use_win) {
      // a staged row starts at the 4-byte aligned address at or below its first pixel
      const int rows = y_hi - y_lo + 1;
      const uintptr_t end = f0 + (size_t)H * W * 3;
      for (int idx = tid; idx < rows * wpitch; idx += (int)blockDim.x) {
        const int r = idx / wpitch, wd = idx - r * wpitch;
        const uintptr_t g0 = f0 + ((size_t)(y_lo + r) * W + x_lo) * 3;
        const uintptr_t ga = (g0 & ~(uintptr_t)3) + (uintptr_t)wd * 4;
        uint32_t v = 0;
        if (ga + 4 <= end) {
          v = __ldg(reinterpret_cast<const uint32_t*>(ga));
        } else {
          for (int k = 0; k < 4; ++k)
            if (ga + k < end) v |= (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(ga + k)) << (8 * k);
        }
        s_win[idx] = v;
      }
      __syncthreads();
    }
    // address of pixel 0 of source row y: inside the window, row (y - y_lo) starts at the aligned
    // word at or below its first pixel x_lo, i.e. `shift` bytes before it
    auto refresh = [&](int y, float2 (&h)[3]) {
      if (use_win) {
        const int shift = (int)((f0 + ((size_t)y * W + x_lo) * 3) & 3);
        load_h<true>(wbase + (size_t)(y - y_lo) * wpitch * 4 + shift - x_lo * 3, oa0, oa1, ob0, ob1,
                     gx2, fx2, h);
      } else {
        load_h<false>(frame + (size_t)y * W * 3, oa0, oa1, ob0, ob1, gx2, fx2, h);
      }
    };
    if (active) {
#pragma unroll 1
      for (int pr = py; pr < py + band; ++pr) {
        uint8_t* drow = dst + (size_t)pr * grid_p * kPatchK * 2;
#pragma unroll 4
        for (int ky = 0; ky < kPatch; ++ky) {
          const int i = pr * kPatch + ky;
          const int y0 = s_y0[i];
          const float fy = s_fy[i];
          const float gy = __fsub_rn(1.0f, fy);
          const int y1 = min(y0 + 1, H - 1);
          if (y0 != top_y) {                                 // CTA-uniform branches
            if (y0 == bot_y) { ht[0] = hb[0]; ht[1] = hb[1]; ht[2] = hb[2]; }
            else refresh(y0, ht);
            top_y = y0;
          }
          if (y1 != bot_y) {
            if (y1 == top_y) { hb[0] = ht[0]; hb[1] = ht[1]; hb[2] = ht[2]; }
            else refresh(y1, hb);
            bot_y = y1;
          }
          const float2 gy2 = make_float2(gy, gy), fy2 = make_float2(fy, fy);
#pragma unroll
          for (int c = 0; c < 3; ++c)
            *reinterpret_cast<uint32_t*>(drow + (c * 256 + ky * kPatch) * 2) =
                blend2(gy2, fy2, ht[c], hb[c], f16);
        }
      }
    }
    py += band;
    if (py < grid_p) __syncthreads();                        // the window is restaged for the next band
  }
}

template <bool kFromState>
static int launch_crop(const uint8_t* frame, int H, int W, const float* src, int n, int ld, float w0,
                       float h0, void* out, bool f16, int S, cudaStream_t stream) {
  static bool attr[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 31]) {
    VPF_CUDA(cudaFuncSetAttribute(crop_col_kernel<kFromState>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * kMaxCrop * 4 + kWinWordsMax * 4));
    attr[dev & 31] = true;
  }
  // One warp per 64 output columns.  Window budget per CTA: what a box of the nominal size needs in
  // one piece, with 15 % head-room for the scale state, capped so that ~16 warps stay resident
  // per SM whatever the box and the crop size (larger windows are staged in bands of patch rows).
  // Explicit boxes (size unknown on the host; one launch per tracker init) get the cap.
  const int warps = (S + 63) / 64;
  const int ctas_per_sm = (16 + warps - 1) / warps;
  int cap = (227 * 1024 / ctas_per_sm - 1024 - 2 * S * 4) / 4;
  if (cap > kWinWordsMax) cap = kWinWordsMax;
  int win_words = cap;
  if (kFromState) {
    const float rows_f = h0 * 1.15f + 4.0f, words_f = (w0 * 1.15f + 4.0f) * 0.75f + 3.0f;
    const float need = rows_f * words_f;
    win_words = need > (float)cap ? cap : (need < 1024.f ? 1024 : (int)need);
  }
  const size_t smem = (size_t)2 * S * 4 + (size_t)win_words * 4;
  VPF_CUDA(launch_pdl(pdl_for_rows((long long)n * (S / kPatch) * (S / kPatch)),
                      crop_col_kernel<kFromState>, dim3(n), dim3(warps * 32), smem, stream, frame,
                      H, W, src, n, ld, w0, h0, (__nv_bfloat16*)out, f16, S, win_words));
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
