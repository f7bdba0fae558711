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

// ---- Box-Muller on 24-bit uniforms (oracle/philox.py box_muller, D14) ------------------------
// The transcendentals run in fp64 and the product is rounded once to fp32: an fp64 libm result
// rounds to the same fp32 value whichever library produced it, so the normals -- and with them
// the whole motion model -- equal the CPU oracle's bit for bit.  6 normals per particle per
// frame: ~25 k fp64 transcendentals at N = 4096, nothing beside the encoder.
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const double u1 = ((double)(ra >> 8) + 1.0) * 5.9604644775390625e-08;  // 2^-24
  const double u2 = (double)(rb >> 8) * 5.9604644775390625e-08;
  const double rad = sqrt(-2.0 * log(u1));
  const double ang = 6.283185307179586 * u2;
  double sn, cs;
  sincos(ang, &sn, &cs);
  z0 = (float)(rad * cs);
  z1 = (float)(rad * sn);
}

// fp32 -> fp32 exponential evaluated in fp64, rounded once (oracle/pf.py exp_f32, D14)
__device__ __forceinline__ float exp_f32(float x) { return (float)exp((double)x); }

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
// fp64 accumulation, one rounding to fp32 (oracle/pf.py weigh): fp64 sums in another order differ
// by ~1e-16 relative, which the rounding absorbs, so log-weights equal the oracle's bit for bit
// given the embeddings.  3 M fp64 FMAs at N = 4096, D = 768.
__global__ void __launch_bounds__(256) loglik_kernel(const float* __restrict__ emb, int D,
                                                      const float* __restrict__ tmpl, int n,
                                                      float tau, float* __restrict__ logw) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float4* e4 = reinterpret_cast<const float4*>(emb + (size_t)warp * D);
  const float4* t4 = reinterpret_cast<const float4*>(tmpl);
  double dot = 0.0, ne = 0.0, nt = 0.0;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 e = e4[i];
    const float4 t = __ldg(t4 + i);
    const double ex = e.x, ey = e.y, ez = e.z, ew = e.w, tx = t.x, ty = t.y, tz = t.z, tw = t.w;
    dot += ex * tx + ey * ty + ez * tz + ew * tw;
    ne += ex * ex + ey * ey + ez * ez + ew * ew;
    nt += tx * tx + ty * ty + tz * tz + tw * tw;
  }
  dot = warp_sum(dot);
  ne = warp_sum(ne);
  nt = warp_sum(nt);
  if (lane == 0) {
    const float c = (float)(dot / (fmax(sqrt(ne), 1e-8) * fmax(sqrt(nt), 1e-8)));
    logw[warp] = __fdiv_rn(c, tau);
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
    float e = exp_f32(__fsub_rn(logw[i], m));
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
    double* __restrict__ stats, int accumulate) {
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
      float w = exp_f32(__fsub_rn(logw[i], lse));
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
    stats[tid] = accumulate ? stats[tid] + a : a;
  }
}

// ---- systematic resampling + gather --------------------------------------------------------
// Where the R ranks' pieces of the global CDF / state live.  Two layouts feed the same kernel:
//   separate arrays (single GPU, tests):  cdf f64 [R][n], totals f64 [R], state f32 [R][6][n];
//   packed records (multi-GPU exchange):  rank r's record at rec + r*stride (f64 slots):
//     [0] total  [1..6) stats  [8, 8+n_r) local inclusive CDF  [8+n_max, ...) state f32 [6][n_max]
//   read in place from the all-gather's receive buffer -- no repacking kernels.
// Shards may be uneven: shard_off[r] = first global particle id of rank r (device, R+1 ints);
// nullptr = equal shards of n_equal.
struct ResampleSrc {
  const double* cdf;    long long cdf_stride;
  const double* tot;    long long tot_stride;
  const float* st;      long long st_stride;   int st_ld;
  const int32_t* shard_off;                    int n_equal;
  const double* stats;  long long stats_stride;  double* stats_out;   // optional rank reduction
};

static constexpr int kMaxRanks = 64;

__global__ void __launch_bounds__(256) resample_kernel(
    ResampleSrc src, int R, int n_local, uint32_t id_offset, double u,
    const uint32_t* __restrict__ frame_idx_dev, uint32_t k0, uint32_t k1, uint32_t stream_id,
    float* __restrict__ state_out, int ld_out, int32_t* __restrict__ idx) {
  __shared__ double offs[kMaxRanks + 1];   // exclusive prefix of rank totals; offs[R] = C_last
  __shared__ int soff[kMaxRanks + 1];      // first global id of each rank; soff[R] = N
  if (frame_idx_dev) {              // CUDA-graph replay: derive the frame's uniform on device
    Philox4 r = philox4x32_10(0xFFFFFFFFu, *frame_idx_dev, stream_id, 2u, k0, k1);
    const uint64_t v = (((uint64_t)r.v[0] << 21) ^ ((uint64_t)r.v[1] >> 11)) & ((1ull << 53) - 1);
    u = (double)v * (1.0 / 9007199254740992.0);
  }
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int r = 0; r < R; ++r) { offs[r] = a; a = __dadd_rn(a, src.tot[r * src.tot_stride]); }
    offs[R] = a;                   // exact: every total is a multiple of 2^-40, the sum is < 2
  }
  if (threadIdx.x <= R)
    soff[threadIdx.x] = src.shard_off ? src.shard_off[threadIdx.x] : (int)threadIdx.x * src.n_equal;
  __syncthreads();
  if (src.stats_out && blockIdx.x == 0 && threadIdx.x < 5) {
    double a = 0.0;                // fixed rank order: every rank obtains the same estimate
    for (int r = 0; r < R; ++r) a += src.stats[r * src.stats_stride + threadIdx.x];
    src.stats_out[threadIdx.x] = a;
  }
  const int N = soff[R];
  const double step = __ddiv_rn(offs[R], (double)N);
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_local) return;
  const double p = __dmul_rn(__dadd_rn((double)(id_offset + (uint32_t)j), u), step);
  // a = #{g : C[g] <= p}, C[g] = offs[r] + cdf_r[i] nondecreasing in g: every rank whose last
  // value offs[r+1] is <= p counts in full; the first rank r with offs[r+1] > p is searched.
  int r = 0;
  while (r < R - 1 && offs[r + 1] <= p) ++r;
  const double* c = src.cdf + r * src.cdf_stride;
  const double base = offs[r];
  int lo = 0, hi = soff[r + 1] - soff[r];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__dadd_rn(base, c[mid]) <= p) lo = mid + 1; else hi = mid;
  }
  int a = soff[r] + lo;
  if (a > N - 1) a = N - 1;
  idx[j] = a;
  const int i = a - soff[r];
  const float* s = src.st + r * src.st_stride + i;
#pragma unroll
  for (int row = 0; row < 6; ++row)
    state_out[(size_t)row * ld_out + j] = s[(size_t)row * src.st_ld];
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
  // terms <= 1 on a 2^-36 grid: the fp64 sum over all particles (of every rank) is exact, hence
  // independent of order and of the number of ranks, only while log2(N_global) + 36 <= 53.  The
  // caller owns the global count (Tracker checks it); the local count is checked here.
  VPF_REQUIRE(n <= (1 << 17), "vpf_sumexp: %d particles exceed 2^17, the bound of the exact fp64 "
              "exp-sum", n);
  sumexp_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, rank_max, R, local_sum);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_weights(const float* logw, int n, const float* rank_max, const double* rank_sum, int R,
                const float* state, int ld, float* w, double* cdf_local, double* total_local,
                double* stats, int accumulate, void* stream) {
  VPF_REQUIRE(logw && rank_max && rank_sum && state && w && cdf_local && total_local && stats &&
                  n > 0 && R > 0 && ld >= n,
              "vpf_weights: bad arguments");
  weights_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, rank_max, rank_sum, R, state, ld,
                                                       w, cdf_local, total_local, stats, accumulate);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

static ResampleSrc arrays_src(const double* cdf_all, const double* totals, const float* state_all,
                              int n_local) {
  ResampleSrc s = {};
  s.cdf = cdf_all;  s.cdf_stride = n_local;
  s.tot = totals;   s.tot_stride = 1;
  s.st = state_all; s.st_stride = 6ll * n_local;  s.st_ld = n_local;
  s.shard_off = nullptr;  s.n_equal = n_local;
  return s;
}

int vpf_resample(const double* cdf_all, const double* totals, int R, int n_local,
                 uint32_t id_offset, double u, const float* state_all, float* state_out,
                 int ld_out, int32_t* idx, void* stream) {
  VPF_REQUIRE(cdf_all && totals && state_all && state_out && idx && R > 0 && R <= kMaxRanks &&
                  n_local > 0 && ld_out >= n_local && u >= 0.0 && u < 1.0,
              "vpf_resample: bad arguments (need 0 <= u < 1, R <= 64)");
  VPF_REQUIRE((long long)R * n_local < (1ll << 31), "vpf_resample: too many particles");
  resample_kernel<<<(n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      arrays_src(cdf_all, totals, state_all, n_local), R, n_local, id_offset, u, nullptr, 0u, 0u,
      0u, state_out, ld_out, idx);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_resample_dev(const double* cdf_all, const double* totals, int R, int n_local,
                     uint32_t id_offset, uint64_t seed, uint32_t stream_id,
                     const uint32_t* frame_idx_dev, const float* state_all, float* state_out,
                     int ld_out, int32_t* idx, void* stream) {
  VPF_REQUIRE(cdf_all && totals && state_all && state_out && idx && frame_idx_dev && R > 0 &&
                  R <= kMaxRanks && n_local > 0 && ld_out >= n_local,
              "vpf_resample_dev: bad arguments");
  VPF_REQUIRE((long long)R * n_local < (1ll << 31), "vpf_resample_dev: too many particles");
  uint32_t k0, k1; split_seed(seed, k0, k1);
  resample_kernel<<<(n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      arrays_src(cdf_all, totals, state_all, n_local), R, n_local, id_offset, 0.0, frame_idx_dev,
      k0, k1, stream_id, state_out, ld_out, idx);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

int vpf_resample_records(const double* records, long long rec_stride, int n_max,
                         const int32_t* shard_off, int R, int n_local, uint32_t id_offset, double u,
                         const uint32_t* frame_idx_dev, uint64_t seed, uint32_t stream_id,
                         float* state_out, int ld_out, int32_t* idx, double* stats_out,
                         void* stream) {
  VPF_REQUIRE(records && shard_off && state_out && idx && R > 0 && R <= kMaxRanks && n_local > 0 &&
                  n_max >= n_local && ld_out >= n_local &&
                  rec_stride >= VPF_RECORD_HEADER + 4ll * n_max,
              "vpf_resample_records: bad arguments (R <= 64, rec_stride >= 8 + 4 n_max)");
  VPF_REQUIRE(frame_idx_dev || (u >= 0.0 && u < 1.0), "vpf_resample_records: need 0 <= u < 1");
  ResampleSrc s = {};
  s.cdf = records + VPF_RECORD_HEADER;  s.cdf_stride = rec_stride;
  s.tot = records;                      s.tot_stride = rec_stride;
  s.st = reinterpret_cast<const float*>(records + VPF_RECORD_HEADER + n_max);
  s.st_stride = 2 * rec_stride;         s.st_ld = n_max;
  s.shard_off = shard_off;              s.n_equal = 0;
  s.stats = records + 1;                s.stats_stride = rec_stride;  s.stats_out = stats_out;
  uint32_t k0, k1; split_seed(seed, k0, k1);
  resample_kernel<<<(n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      s, R, n_local, id_offset, u, frame_idx_dev, k0, k1, stream_id, state_out, ld_out, idx);
  VPF_LAUNCH_CHECK();
  return VPF_OK;
}

}  // extern "C"
