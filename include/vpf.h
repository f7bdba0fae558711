/*
 * vpf.h -- C ABI of libvpf.so: the B200 (sm_100a) implementation of the ViT particle-filter
 * tracking hot path.
 *
 * Drop-in boundary.  The upstream project (tugitbartlomiej/ViTParticleFilterTracker) defines no
 * FFI, plugin or operator interface: its tree is a single README.md.  The only contract it
 * states is README.md:42 ("provide the input data (e.g., video feed or images) ... The system
 * will output the tracked positions of the objects in real-time"), README.md:7 (ViT feature
 * extraction) and README.md:8 (particle filter state estimation).  Each entry point below
 * therefore cites the README line whose function it implements and the oracle function
 * (oracle/) that defines its result.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, a `void* stream` that is a cudaStream_t
 *     (torch.cuda.current_stream().cuda_stream); no torch types.
 *   - every tensor is allocated and owned by the caller; the library never frees caller memory
 *     and never synchronises: all work is enqueued on `stream` (CUDA-graph capturable).
 *   - every entry returns 0 on success or a negative vpf_status; a message for the calling
 *     thread's last failure is available from vpf_last_error().
 *   - particle state is float32 structure-of-arrays [6][ld]: rows cx, cy, vx, vy, s, spare
 *     (`ld` >= n is the row pitch in elements).
 *   - "rank" arguments serve the particle-sharded multi-GPU path: rank r of R owns global
 *     particle ids [r*n_local, (r+1)*n_local).  Single-GPU callers pass R = 1, id_offset = 0.
 */
#ifndef VPF_H_
#define VPF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPF_ABI_VERSION 7

/* 16-bit operand type of crops / ViT activations / weights (fp32 accumulation either way) */
#define VPF_DTYPE_BF16 0
#define VPF_DTYPE_F16 1
/* storage type of the encoder's residual stream X [n*T][D]: fp32 (default, what the parity
 * figures are quoted on) or fp16 with saturating conversion (half the residual bytes: X is 35 %
 * of a block's HBM traffic; adds ~1e-3 relative error, |x| must stay below 65504) */
#define VPF_RESIDUAL_F32 0
#define VPF_RESIDUAL_F16 1

typedef enum {
  VPF_OK = 0,
  VPF_ERR_INVALID = -1,   /* bad argument */
  VPF_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  VPF_ERR_STATE = -3,     /* handle not ready (weights / workspace missing) */
  VPF_ERR_UNSUPPORTED = -4
} vpf_status;

typedef struct vpf_vit vpf_vit; /* opaque ViT encoder handle */

typedef struct {
  int32_t dim;       /* D: 192 / 384 / 768 / 1024 (multiple of 64) */
  int32_t depth;     /* L */
  int32_t heads;     /* H, D / H must be 64 */
  int32_t max_crops; /* capacity: largest n_crops of one vpf_vit_forward call */
  int32_t dtype;     /* VPF_DTYPE_BF16 (default) or VPF_DTYPE_F16 */
  int32_t crop_size; /* S: input crop size in pixels, 0 or 224 = the default; a smaller multiple
                        of 16 gives P = (S/16)^2 patches and T = P + 1 tokens per crop (SURVEY.md
                        section 8f-2 "native-resolution crops"; oracle/vit.py resamples pos_embed) */
  int32_t residual_dtype; /* VPF_RESIDUAL_F32 (default) or VPF_RESIDUAL_F16 */
} vpf_vit_config;

/* ---- library ------------------------------------------------------------------------------ */
int vpf_abi_version(void);
const char* vpf_last_error(void);

/* ---- a3: motion model (README.md:8 "Particle Filter"; oracle/pf.py propagate, D7/D8) -------
 * Philox4x32-10, key = seed, counter = (id_offset + i, frame_idx, stream_id, draw).          */
int vpf_init_particles(float* state, int n, int ld, float cx, float cy, float sigma_init,
                       uint64_t seed, uint32_t stream_id, uint32_t id_offset, void* stream);
int vpf_propagate(float* state, int n, int ld, float sigma_pos, float sigma_vel,
                  float sigma_scale, uint64_t seed, uint32_t frame_idx, uint32_t stream_id,
                  uint32_t id_offset, void* stream);
/* CUDA-graph variants: the frame counter lives in device memory (advanced by vpf_frame_advance),
 * so one captured graph can be replayed for every frame.  Same arithmetic as the host-argument
 * forms; vpf_resample_dev derives the frame's uniform from (seed, *frame_idx_dev, stream_id).  */
int vpf_frame_advance(uint32_t* frame_idx_dev, void* stream);
int vpf_propagate_dev(float* state, int n, int ld, float sigma_pos, float sigma_vel,
                      float sigma_scale, uint64_t seed, const uint32_t* frame_idx_dev,
                      uint32_t stream_id, uint32_t id_offset, void* stream);
/* test hooks: raw Philox words [4][n] for counters (id_offset+i, c1, c2, c3); the six normals
 * [6][n] that propagate consumes; the fp64 resampling uniform (host, no device work).        */
int vpf_philox_fill(uint32_t* out, int n, uint32_t id_offset, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint64_t seed, void* stream);
int vpf_normals_fill(float* out, int n, uint32_t id_offset, uint32_t frame_idx,
                     uint32_t stream_id, uint64_t seed, void* stream);
double vpf_resample_uniform(uint64_t seed, uint32_t frame_idx, uint32_t stream_id);

/* ---- a4: crop + bilinear resize + normalise + patchify (README.md:7; oracle/crop.py, D3/D4) -
 * frame: uint8 [H][W][3]; boxes come either from the SoA state (w = w0*exp(s), h = h0*exp(s))
 * or from an explicit [n][4] (cx,cy,w,h) array.  crop_size S: 224, or a smaller multiple of 16.
 * out: bf16 or fp16 (`dtype`) [n*(S/16)^2][768], k = c*256+ky*16+kx.                           */
int vpf_crop_state(const uint8_t* frame, int H, int W, const float* state, int n, int ld,
                   float w0, float h0, void* out16, int dtype, int crop_size, void* stream);
int vpf_crop_boxes(const uint8_t* frame, int H, int W, const float* boxes, int n,
                   void* out16, int dtype, int crop_size, void* stream);

/* ---- a5: ViT/16 encoder (README.md:7; oracle/vit.py, D1/D2) --------------------------------
 * weights: array of VPF_VIT_N_GLOBAL + 12*depth + 2 device pointers in this order:
 *   patch_w bf16[D][768], patch_b f32[D], cls f32[D], pos f32[T][D] (T = 197 at crop size 224),
 *   per block: ln1_g, ln1_b f32[D]; qkv_w bf16[3D][D], qkv_b f32[3D]; proj_w bf16[D][D],
 *              proj_b f32[D]; ln2_g, ln2_b f32[D]; fc1_w bf16[4D][D], fc1_b f32[4D];
 *              fc2_w bf16[D][4D], fc2_b f32[D];
 *   norm_g, norm_b f32[D].
 * (Linear weights keep PyTorch's [out][in] layout: that is the K-major B operand tcgen05 wants.)
 * workspace: caller-allocated, >= vpf_vit_workspace_bytes(cfg), 1024-byte aligned.            */
#define VPF_VIT_N_GLOBAL 4
int vpf_vit_create(vpf_vit** out, const vpf_vit_config* cfg);
int vpf_vit_destroy(vpf_vit* h);
size_t vpf_vit_workspace_bytes(const vpf_vit_config* cfg);
int vpf_vit_set_weights(vpf_vit* h, const void* const* dev_ptrs, int n_ptrs);
int vpf_vit_set_workspace(vpf_vit* h, void* workspace, size_t bytes);
/* patches: bf16 [n_crops*P][768] (output of vpf_crop_*, P = 196 at crop size 224);
 * emb: f32 [n_crops][D].                                                                      */
int vpf_vit_forward(vpf_vit* h, const void* patches_bf16, int n_crops, float* emb, void* stream);
/* test hooks for single stages: C[M][N] = A[M][K] * W[N][K]^T (+ epilogue `mode`), LayerNorm,
 * attention.  mode: 0 bias->bf16, 1 bias+GELU->bf16, 2 bias+residual f32 in place,
 * 3 patch-embed (bias + pos, row remap into the (patches_per_crop + 1)-token stream, f32).
 * residual_dtype (modes 2, 3): VPF_RESIDUAL_F32, or VPF_RESIDUAL_F16 when `out` is an fp16 stream. */
int vpf_gemm_bf16(const void* A, const void* W, const float* bias, void* out, const float* pos,
                  int M, int N, int K, int mode, int patches_per_crop, int residual_dtype,
                  void* stream);
int vpf_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16,
                  int rows, int D, void* stream);
/* qkv: bf16 [n_crops*tokens][3*64*heads], out: bf16 [n_crops*tokens][64*heads]; 2 <= tokens <= 197 */
int vpf_attention(const void* qkv_bf16, void* out_bf16, int n_crops, int heads, int tokens,
                  void* stream);
/* LayerNorm folded into the GEMMs that consume it (fp16 residual stream only; README.md:7; the
 * arithmetic of oracle/vit.py's norm1 -> qkv and norm2 -> fc1, reassociated):
 *     LN(x) W^T + b  =  r (x W'^T) + b',   W'_jk = gamma_k W_jk - mean_k(gamma_k W_jk),  b' = b + W beta
 * (r = 1/sqrt(var(x) + eps); subtracting each weight row's mean over k makes x W'^T equal to
 * (x - mean(x)) (gamma (.) W)^T, so the row mean is never needed).  The fp16 stream itself is the A
 * operand and the normalised activation never touches HBM.
 * dev_ptrs: 4 per block, in block order: qkv_wf fp16 [3D][D], qkv_bf f32 [3D], fc1_wf fp16 [4D][D],
 * fc1_bf f32 [4D].
 * W' is fp16 whatever cfg.dtype says: both MMA operands must share one 16-bit type and A is the
 * fp16 stream; the GEMMs' outputs stay in cfg.dtype.
 * The row statistics (mu, r) come from the residual GEMMs' epilogues as per-column-group partial
 * sums (env VPF_LN_FOLD=kernel: from a row-statistics kernel; VPF_LN_FOLD=0: folding ignored).
 * vpf_vit_ln_fold: 0 = not folded, 1 = statistics kernel, 2 = epilogue sums.                     */
int vpf_vit_set_ln_fold(vpf_vit* h, const void* const* dev_ptrs, int n_ptrs);
int vpf_vit_ln_fold(const vpf_vit* h);
/* test hooks of the folded path: column groups a D-wide residual GEMM splits a row into; final
 * row statistics stats[rows][2] = (mean, 1/sqrt(var + 1e-6)) of an fp16 [rows][D] matrix; the
 * folded GEMM (X16, Wf fp16; out in `dtype`; stats_parts = 0: stats = [M][2] final; > 0: [stats_parts][M][2] partial (sum x,
 * sum x^2)); the fp16-stream residual GEMM that also writes those partial sums
 * (stats_out f32 [vpf_ln_stats_parts(N)][M][2]).                                                  */
int vpf_ln_stats_parts(int D);
int vpf_rowstats16(const void* x16, float* stats, int rows, int D, void* stream);
int vpf_gemm_ln(const void* X16, const void* Wf, const float* bias_f, const float* stats,
                int stats_parts, void* out, int M, int N, int K, int gelu, int dtype, void* stream);
int vpf_gemm_residual_stats(const void* A, const void* W, const float* bias, void* x16,
                            float* stats_out, int M, int N, int K, int dtype, void* stream);
/* number of kernels one vpf_vit_forward enqueues (for bench.py's gpu_launches claim).         */
int vpf_vit_launches(const vpf_vit* h);
/* Measurement aid (SURVEY.md section 8d "per-stage times"): with timing enabled every launch of a
 * forward is bracketed by CUDA events on the caller's stream (not capturable in a graph while on);
 * vpf_vit_timing waits for the last forward and returns its device milliseconds summed per kind:
 * [0] patch embed  [1] LayerNorm  [2] qkv GEMM  [3] attention  [4] proj GEMM  [5] fc1 GEMM
 * [6] fc2 GEMM  [7] last block (CLS-only) + final norm.  n_kinds >= VPF_VIT_TIMING_KINDS.        */
#define VPF_VIT_TIMING_KINDS 8
int vpf_vit_set_timing(vpf_vit* h, int enable);
int vpf_vit_timing(vpf_vit* h, float* ms_by_kind, int n_kinds);

/* ---- a6: likelihood, normalisation, ESS, estimate (README.md:8; oracle/pf.py weigh, D9/D10) -
 * vpf_loglik:   logw[i] = cos(emb[i], tmpl) / tau;  local_max[0] = max_i logw[i].
 * vpf_sumexp:   m = max(rank_max[0..R)); local_sum[0] = sum_i snap36(exp(logw[i] - m)) (fp64,
 *               exact -> independent of summation order and of R).
 * vpf_weights:  lse = float(m + log(sum_r rank_sum[r])); w[i] = exp(logw[i] - lse);
 *               cdf_local[i] = inclusive fp64 scan of snap40(w) (exact), total_local[0] = last;
 *               stats[0..5) (+)= (sum w, sum w^2, sum w*cx, sum w*cy, sum w*s) in fp64:
 *               accumulate != 0 adds to what is there (several shards into one vector, the
 *               caller zeroes it first), accumulate == 0 overwrites.
 * All exponentials are exp() in fp64 of an fp32 difference, rounded once to fp32, and the
 * likelihood accumulates in fp64: given the embeddings, log-weights, weights
 * and CDF equal the CPU oracle's bit for bit.  N_global <= 2^17 (exactness of the exp-sum).   */
int vpf_loglik(const float* emb, int D, const float* tmpl, int n, float tau, float* logw,
               float* local_max, void* stream);
int vpf_sumexp(const float* logw, int n, const float* rank_max, int R, double* local_sum,
               void* stream);
int vpf_weights(const float* logw, int n, const float* rank_max, const double* rank_sum, int R,
                const float* state, int ld, float* w, double* cdf_local, double* total_local,
                double* stats, int accumulate, void* stream);

/* ---- a7 + a8: systematic resampling + state gather (README.md:8; oracle/pf.py, D6) ---------
 * Global CDF C[r*n_local + i] = sum_{r'<r} totals[r'] + cdf_all[r][i]  (exact in fp64).
 * For local output slot j (global slot id_offset + j): p = (id_offset + j + u) * (C_last / N),
 * a = #{g : C[g] <= p} clamped to N-1, idx[j] = a, state_out[:, j] = state_all[a / n_local][:,
 * a % n_local].  cdf_all: f64 [R][n_local]; totals: f64 [R]; state_all: f32 [R][6][n_local].   */
int vpf_resample(const double* cdf_all, const double* totals, int R, int n_local,
                 uint32_t id_offset, double u, const float* state_all, float* state_out,
                 int ld_out, int32_t* idx, void* stream);

int vpf_resample_dev(const double* cdf_all, const double* totals, int R, int n_local,
                     uint32_t id_offset, uint64_t seed, uint32_t stream_id,
                     const uint32_t* frame_idx_dev, const float* state_all, float* state_out,
                     int ld_out, int32_t* idx, void* stream);


/* Multi-GPU form: the R ranks' pieces are read in place from the
 * all-gather's receive buffer.  Rank r's record starts at records + r*rec_stride (f64 slots):
 *   [0] total   [1..6) stats (vpf_weights' five sums)   [6..8) unused
 *   [8, 8+n_r) local inclusive CDF   [8+n_max, 8+4*n_max) state f32 [6][n_max] (ld = n_max)
 * Shards may be uneven: shard_off (device, R+1 int32) holds each rank's first global particle id,
 * shard_off[R] = N.  This rank resamples its n_local output slots [id_offset, id_offset+n_local).
 * stats_out (optional, 5 f64) receives the sum over ranks of the records' stats, in rank order.
 * frame_idx_dev != NULL: the uniform is derived on the device (CUDA-graph replay), `u` ignored. */
#define VPF_RECORD_HEADER 8
int vpf_resample_records(const double* records, long long rec_stride, int n_max,
                         const int32_t* shard_off, int R, int n_local, uint32_t id_offset, double u,
                         const uint32_t* frame_idx_dev, uint64_t seed, uint32_t stream_id,
                         float* state_out, int ld_out, int32_t* idx, double* stats_out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPF_H_ */
