// Particle-filter kernels: Philox motion model, likelihood, exact fp64 normalisation and CDF,
// systematic resampling + state gather.  Oracle: oracle/pf.py, oracle/philox.py
// (upstream anchor README.md:8 "Particle Filter").
//
// All of these are latency-bound (<= 1 MB of traffic at N = 32768): they are written for few
// launches, coalesced SoA access and *order-independent* arithmetic (weights and exp terms are
// snapped to a power-of-two grid so that fp64 sums are exact), which is what makes resampling
// indices bit-exact against the oracle and independent of the number of ranks.
#include "common.cuh"

#include <math.h>

namespace vpf {

static constexpr double kGrid40 = 1099511627776.0;           // 2^40
static constexpr double kInvGrid40 = 1.0 / 1099511627776.0;  // 2^-40
static constexpr double kGrid36 = 68719476736.0;             // 2^36
static constexpr double kInvGrid36 = 1.0 / 68719476736.0;

// ---- Box-Muller on 24-bit uniforms (oracle/philox.py box_muller) --------------------------
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  float u1 = __fmul_rn(__fadd_rn((float)(ra >> 8), 1.0f), 5.9604644775390625e-08f);  // 2^-24
  float u2 = __fmul_rn((float)(rb >> 8), 5.9604644775390625e-08f);
  float rad = sqrtf(__fmul_rn(-2.0f, logf(u1)));
  float ang = __fmul_rn(6.2831855f, u2);
  z0 = __fmul_rn(rad, cosf(ang));
  z1 = __fmul_rn(rad, sinf(ang));
}

__device__ __forceinline__ void six_normals(uint32_t id, uint32_t frame_idx, uint32_t stream_id,
                                            uint32_t k0, uint32_t k1, float z[6]) {
  Philox4 a = philox4x32_10(id, frame_idx, stream_id, 0u, k0, k1);
  Philox4 b = philox4x32_10(id, frame_idx, stream_id, 1u, k0, k1);
  box_muller(a.v[0], a.v[1], z[0], z[1]);
  box_muller(a.v[2], a.v[3], z[2], z[3]);
  box_muller(b.v[0], b.v[1], z[4], z[5]);
}

__global__ void philox_fill_kernel(uint32_t* __restrict__ out, int n, uint32_t id_offset,
                                   uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                   uint32_t k1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 r = philox4x32_10(id_offset + (uint32_t)i, c1, c2, c3, k0, k1);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[(size_t)j * n + i] = r.v[j];
}

__global__ void normals_fill_kernel(float* __restrict__ out, int n, uint32_t id_offset,
                                    uint32_t frame_idx, uint32_t stream_id, uint32_t k0,
                                    uint32_t k1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z[6];
  six_normals(id_offset + (uint32_t)i, frame_idx, stream_id, k0, k1, z);
#pragma unroll
  for (int j = 0; j < 6; ++j) out[(size_t)j * n + i] = z[j];
}

__global__ void init_particles_kernel(float* __restrict__ st, int n, int ld, float cx, float cy,
                                      float sigma_init, uint32_t stream_id, uint32_t id_offset,
                                      uint32_t k0, uint32_t k1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z[6];
  six_normals(id_offset + (uint32_t)i, 0u, stream_id, k0, k1, z);
  st[0 * (size_t)ld + i] = __fadd_rn(cx, __fmul_rn(sigma_init, z[0]));
  st[1 * (size_t)ld + i] = __fadd_rn(cy, __fmul_rn(sigma_init, z[1]));
#pragma unroll
  for (int r = 2; r < 6; ++r) st[r * (size_t)ld + i] = 0.0f;
}

// constant velocity, dt = 1: positions use the velocity from before this step's velocity noise
__global__ void propagate_kernel(float* __restrict__ st, int n, int ld, float sp, float sv,
                                 float ss, uint32_t frame_idx,
                                 const uint32_t* __restrict__ frame_idx_dev, uint32_t stream_id,
                                 uint32_t id_offset, uint32_t k0, uint32_t k1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (frame_idx_dev) frame_idx = *frame_idx_dev;   // CUDA-graph replay: the counter lives on device
  float z[6];
  six_normals(id_offset + (uint32_t)i, frame_idx, stream_id, k0, k1, z);
  float cx = st[0 * (size_t)ld + i], cy = st[1 * (size_t)ld + i];
  float vx = st[2 * (size_t)ld + i], vy = st[3 * (size_t)ld + i];
  float s = st[4 * (size_t)ld + i];
  st[0 * (size_t)ld + i] = __fadd_rn(__fadd_rn(cx, vx), __fmul_rn(sp, z[0]));
  st[1 * (size_t)ld + i] = __fadd_rn(__fadd_rn(cy, vy), __fmul_rn(sp, z[1]));
  st[2 * (size_t)ld + i] = __fadd_rn(vx, __fmul_rn(sv, z[2]));
  st[3 * (size_t)ld + i] = __fadd_rn(vy, __fmul_rn(sv, z[3]));
  st[4 * (size_t)ld + i] = __fadd_rn(s, __fmul_rn(ss, z[4]));
}

// ---- likelihood: one warp per particle -----------------------------------------------------
__global__ void __launch_bounds__(256) loglik_kernel(const float* __restrict__ emb, int D,
                                                      const float* __restrict__ tmpl, int n,
                                                      float tau, float* __restrict__ logw) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float4* e4 = reinterpret_cast<const float4*>(emb + (size_t)warp * D);
  const float4* t4 = reinterpret_cast<const float4*>(tmpl);
  float dot = 0.f, ne = 0.f, nt = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    float4 e = e4[i];
    float4 t = __ldg(t4 + i);
    dot += e.x * t.x + e.y * t.y + e.z * t.z + e.w * t.w;
    ne += e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
    nt += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  dot = warp_sum(dot);
  ne = warp_sum(ne);
  nt = warp_sum(nt);
  if (lane == 0) {
    float c = dot / (fmaxf(sqrtf(ne), 1e-8f) * fmaxf(sqrtf(nt), 1e-8f));
    logw[warp] = c / tau;
  }
}

__global__ void __launch_bounds__(1024) max_reduce_kernel(const float* __restrict__ x, int n,
                                                           float* __restrict__ out) {
  __shared__ float sm[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = sm[threadIdx.x];
    m = warp_max(m);
    if (threadIdx.x == 0) out[0] = m;
  }
}

__device__ __forceinline__ float ranks_max(const float* rank_max, int R) {
  float m = rank_max[0];
  for (int r = 1; r < R; ++r) m = fmaxf(m, rank_max[r]);
  return m;
}

// exact fp64 sum of exp terms snapped to 2^-36 (order independent)
__global__ void __launch_bounds__(1024) sumexp_kernel(const float* __restrict__ logw, int n,
                                                       const float* __restrict__ rank_max, int R,
                                                       double* __restrict__ local_sum) {
  __shared__ double sm[32];
  const float m = ranks_max(rank_max, R);
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float e = expf(__fsub_rn(logw[i], m));
    acc += rint((double)e * kGrid36) * kInvGrid36;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sm[threadIdx.x];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) local_sum[0] = acc;
  }
}

// w = exp(logw - lse); exact inclusive fp64 scan of snap40(w); ESS / estimate partial sums.
__global__ void __launch_bounds__(1024) weights_kernel(
    const float* __restrict__ logw, int n, const float* __restrict__ rank_max,
    const double* __restrict__ rank_sum, int R, const float* __restrict__ st, int ld,
    float* __restrict__ w_out, double* __restrict__ cdf, double* __restrict__ total,
    double* __restrict__ stats) {
  __shared__ double warp_tot[32];
  __shared__ double stat_sm[32][5];
  __shared__ double carry_sm;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float m = ranks_max(rank_max, R);
  double ssum = 0.0;
  for (int r = 0; r < R; ++r) ssum += rank_sum[r];   // exact: any order gives the same value
  const float lse = (float)((double)m + log(ssum));
  if (tid == 0) carry_sm = 0.0;
  __syncthreads();
  double s_w = 0, s_w2 = 0, s_x = 0, s_y = 0, s_s = 0;
  for (int base = 0; base < n; base += 1024) {
    int i = base + tid;
    double q = 0.0;
    if (i < n) {
      float w = expf(__fsub_rn(logw[i], lse));
      w_out[i] = w;
      q = rint((double)w * kGrid40) * kInvGrid40;
      double wd = (double)w;
      s_w += wd;
      s_w2 += wd * wd;
      s_x += wd * (double)st[0 * (size_t)ld + i];
      s_y += wd * (double)st[1 * (size_t)ld + i];
      s_s += wd * (double)st[4 * (size_t)ld + i];
    }
    // inclusive warp scan (exact additions)
    double v = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      double t = warp_tot[lane];
      double inc = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      warp_tot[lane] = inc - t;  // exclusive prefix of warp totals (exact)
      if (lane == 31) stat_sm[0][0] = inc;  // tile total, parked until read below
    }
    __syncthreads();
    double carry = carry_sm;
    if (i < n) cdf[i] = carry + warp_tot[wid] + v;
    double tile_total = stat_sm[0][0];
    __syncthreads();
    if (tid == 0) carry_sm = carry + tile_total;
    __syncthreads();
  }
  if (tid == 0) total[0] = carry_sm;
  // ESS / estimate sums (output only; not order-critical)
  s_w = warp_sum(s_w); s_w2 = warp_sum(s_w2); s_x = warp_sum(s_x);
  s_y = warp_sum(s_y); s_s = warp_sum(s_s);
  __syncthreads();
  if (lane == 0) {
    stat_sm[wid][0] = s_w; stat_sm[wid][1] = s_w2; stat_sm[wid][2] = s_x;
    stat_sm[wid][3] = s_y; stat_sm[wid][4] = s_s;
  }
  __syncthreads();
  if (tid < 5) {
    double a = 0.0;
    for (int k = 0; k < 32; ++k) a += stat_sm[k][tid];
    stats[tid] += a;
  }
}

// ---- systematic resampling + gather --------------------------------------------------------
__global__ void __launch_bounds__(256) resample_kernel(
    const double* __restrict__ cdf_all, const double* __restrict__ totals, int R, int n_local,
    uint32_t id_offset, double u, const uint32_t* __restrict__ frame_idx_dev, uint32_t k0,
    uint32_t k1, uint32_t stream_id, const float* __restrict__ state_all,
    float* __restrict__ state_out, int ld_out, int32_t* __restrict__ idx) {
  extern __shared__ double offs[];  // [R] exclusive prefix of rank totals
  if (frame_idx_dev) {              // CUDA-graph replay: derive the frame's uniform on device
    Philox4 r = philox4x32_10(0xFFFFFFFFu, *frame_idx_dev, stream_id, 2u, k0, k1);
    const uint64_t v = (((uint64_t)r.v[0] << 21) ^ ((uint64_t)r.v[1] >> 11)) & ((1ull << 53) - 1);
    u = (double)v * (1.0 / 9007199254740992.0);
  }
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int r = 0; r < R; ++r) { offs[r] = a; a = __dadd_rn(a, totals[r]); }
  }
  __syncthreads();
  const int N = R * n_local;
  const double c_last = __dadd_rn(offs[R - 1], cdf_all[(size_t)N - 1]);
  const double step = __ddiv_rn(c_last, (double)N);
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_local) return;
  const double p = __dmul_rn(__dadd_rn((double)(id_offset + (uint32_t)j), u), step);
  int lo = 0, hi = N;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    double c = __dadd_rn(offs[mid / n_local], cdf_all[mid]);
    if (c <= p) lo = mid + 1; else hi = mid;
  }
  int a = lo < N - 1 ? lo : N - 1;
  idx[j] = a;
  int r = a / n_local, i = a - r * n_local;
  const float* src = state_all + (size_t)r * 6 * n_local + i;
#pragma unroll
  for (int row = 0; row < 6; ++row)
    state_out[(size_t)row * ld_out + j] = src[(size_t)row * n_local];
}

static inline void split_seed(uint64_t seed, uint32_t& k0, uint32_t& k1) {
  k0 = (uint32_t)(seed & 0xFFFFFFFFull);
  k1 = (uint32_t)(seed >> 32);
}

}  // namespace vpf

using namespace vpf;

extern "C" {

int vpf_philox_fill(uint32_t* out, int n, uint32_t id_offset, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint64_t seed, void* stream) {
  VPF_REQUIRE(out && n > 0, "vpf_philox_fill: bad arguments");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  philox_fill_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, n, id_offset, c1, c2,
                                                                        c3, k0, k1);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_normals_fill(float* out, int n, uint32_t id_offset, uint32_t frame_idx,
                     uint32_t stream_id, uint64_t seed, void* stream) {
  VPF_REQUIRE(out && n > 0, "vpf_normals_fill: bad arguments");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  normals_fill_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, n, id_offset,
                                                                         frame_idx, stream_id, k0,
                                                                         k1);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

double vpf_resample_uniform(uint64_t seed, uint32_t frame_idx, uint32_t stream_id) {
  uint32_t k0, k1; split_seed(seed, k0, k1);
  Philox4 r = philox4x32_10(0xFFFFFFFFu, frame_idx, stream_id, 2u, k0, k1);
  uint64_t v = (((uint64_t)r.v[0] << 21) ^ ((uint64_t)r.v[1] >> 11)) & ((1ull << 53) - 1);
  return (double)v * (1.0 / 9007199254740992.0);
}

int vpf_init_particles(float* state, int n, int ld, float cx, float cy, float sigma_init,
                       uint64_t seed, uint32_t stream_id, uint32_t id_offset, void* stream) {
  VPF_REQUIRE(state && n > 0 && ld >= n, "vpf_init_particles: bad arguments");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  init_particles_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      state, n, ld, cx, cy, sigma_init, stream_id, id_offset, k0, k1);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_propagate(float* state, int n, int ld, float sigma_pos, float sigma_vel,
                  float sigma_scale, uint64_t seed, uint32_t frame_idx, uint32_t stream_id,
                  uint32_t id_offset, void* stream) {
  VPF_REQUIRE(state && n > 0 && ld >= n, "vpf_propagate: bad arguments");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  propagate_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      state, n, ld, sigma_pos, sigma_vel, sigma_scale, frame_idx, nullptr, stream_id, id_offset, k0,
      k1);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

__global__ void frame_advance_kernel(uint32_t* f) { *f += 1u; }

int vpf_frame_advance(uint32_t* frame_idx_dev, void* stream) {
  VPF_REQUIRE(frame_idx_dev, "vpf_frame_advance: null argument");
  frame_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(frame_idx_dev);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_propagate_dev(float* state, int n, int ld, float sigma_pos, float sigma_vel,
                      float sigma_scale, uint64_t seed, const uint32_t* frame_idx_dev,
                      uint32_t stream_id, uint32_t id_offset, void* stream) {
  VPF_REQUIRE(state && frame_idx_dev && n > 0 && ld >= n, "vpf_propagate_dev: bad arguments");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  propagate_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      state, n, ld, sigma_pos, sigma_vel, sigma_scale, 0u, frame_idx_dev, stream_id, id_offset, k0,
      k1);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_loglik(const float* emb, int D, const float* tmpl, int n, float tau, float* logw,
               float* local_max, void* stream) {
  VPF_REQUIRE(emb && tmpl && logw && local_max && n > 0 && D > 0 && D % 4 == 0 && tau > 0.f,
              "vpf_loglik: bad arguments (D must be a multiple of 4)");
  cudaStream_t s = (cudaStream_t)stream;
  loglik_kernel<<<(n + 7) / 8, 256, 0, s>>>(emb, D, tmpl, n, tau, logw);
  VPF_LAUNCH_CHECK();
  max_reduce_kernel<<<1, 1024, 0, s>>>(logw, n, local_max);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_sumexp(const float* logw, int n, const float* rank_max, int R, double* local_sum,
               void* stream) {
  VPF_REQUIRE(logw && rank_max && local_sum && n > 0 && R > 0, "vpf_sumexp: bad arguments");
  sumexp_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, rank_max, R, local_sum);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_weights(const float* logw, int n, const float* rank_max, const double* rank_sum, int R,
                const float* state, int ld, float* w, double* cdf_local, double* total_local,
                double* stats, void* stream) {
  VPF_REQUIRE(logw && rank_max && rank_sum && state && w && cdf_local && total_local && stats &&
                  n > 0 && R > 0 && ld >= n,
              "vpf_weights: bad arguments");
  weights_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, rank_max, rank_sum, R, state, ld,
                                                       w, cdf_local, total_local, stats);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_resample(const double* cdf_all, const double* totals, int R, int n_local,
                 uint32_t id_offset, double u, const float* state_all, float* state_out,
                 int ld_out, int32_t* idx, void* stream) {
  VPF_REQUIRE(cdf_all && totals && state_all && state_out && idx && R > 0 && n_local > 0 &&
                  ld_out >= n_local && u >= 0.0 && u < 1.0,
              "vpf_resample: bad arguments (need 0 <= u < 1)");
  VPF_REQUIRE((long long)R * n_local < (1ll << 31), "vpf_resample: too many particles");
  resample_kernel<<<(n_local + 255) / 256, 256, R * sizeof(double), (cudaStream_t)stream>>>(
      cdf_all, totals, R, n_local, id_offset, u, nullptr, 0u, 0u, 0u, state_all, state_out, ld_out,
      idx);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_resample_dev(const double* cdf_all, const double* totals, int R, int n_local,
                     uint32_t id_offset, uint64_t seed, uint32_t stream_id,
                     const uint32_t* frame_idx_dev, const float* state_all, float* state_out,
                     int ld_out, int32_t* idx, void* stream) {
  VPF_REQUIRE(cdf_all && totals && state_all && state_out && idx && frame_idx_dev && R > 0 &&
                  n_local > 0 && ld_out >= n_local,
              "vpf_resample_dev: bad arguments");
  VPF_REQUIRE((long long)R * n_local < (1ll << 31), "vpf_resample_dev: too many particles");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  resample_kernel<<<(n_local + 255) / 256, 256, R * sizeof(double), (cudaStream_t)stream>>>(
      cdf_all, totals, R, n_local, id_offset, 0.0, frame_idx_dev, k0, k1, stream_id, state_all,
      state_out, ld_out, idx);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

}  // extern "C"
