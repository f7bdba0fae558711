// Micro-benchmark (not part of libvpf): tensor-memory read bandwidth (tcgen05.ld 32x32b.x32) per
// SM as a function of the number of reading warps, and whether those reads overlap with
// tcgen05.mma accumulator traffic or serialise with it.  Motivation: the qkv / fc1 GEMMs (K = 768)
// run at 75-80 % tensor-pipe-active while fc2 (K = 3072) runs at 93.5 %; 256/K is exactly the
// ratio a non-overlapped 64 B/clk accumulator read-out would cost.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../vitparticlefiltertracker_b200/csrc/sm100.cuh"

using namespace vpf;

struct Result { long long mma_cycles, ld_cycles; };

// block: warp 0 unused, warp 1 = MMA issuer + TMEM owner, warps 2.. = readers
__global__ void __launch_bounds__(576, 1)
probe(int mma_iters, int ld_iters, int ld_warps, int pipelined, int n_cols, Result* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ long long t_mma, t_ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); t_mma = 0; t_ld = 0; }
  if (warp == 1) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1) {
    if (lane == 0 && mma_iters > 0) {
      const uint32_t idesc = make_idesc_bf16(128, n_cols);
      const uint64_t adesc = make_smem_desc_k_sw128(smem_u32(tiles));
      const uint64_t bdesc = make_smem_desc_k_sw128(smem_u32(tiles + 16384));
      const long long t0 = clock64();
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0, 1);
      t_mma = clock64() - t0;
    }
  } else if (warp >= 2 && warp < 2 + ld_warps && ld_iters > 0) {
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;                       // column group of 64 columns
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)((grp & 3) * 64);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ld_iters; ++it) {
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
      if (pipelined) tmem_ld_32x32(taddr + 32u, vb);
      tmem_ld_wait();
      if (!pipelined) { tmem_ld_32x32(taddr + 32u, vb); tmem_ld_wait(); }
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= va[i] ^ vb[i];
    }
    const long long dt = clock64() - t0;
    if (lane == 0) atomicMax((unsigned long long*)&t_ld, (unsigned long long)dt);
    if (acc == 0x12345u) printf("%u\n", acc);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { out[blockIdx.x].mma_cycles = t_mma; out[blockIdx.x].ld_cycles = t_ld; }
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Result* d;
  cudaMalloc(&d, sizeof(Result) * sms);
  Result* h = (Result*)malloc(sizeof(Result) * sms);
  auto run = [&](const char* name, int mma_iters, int ld_iters, int ld_warps, int pipelined, int n_cols) {
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<sms, 576, smem>>>(mma_iters, ld_iters, ld_warps, pipelined, n_cols, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, sizeof(Result) * sms, cudaMemcpyDeviceToHost);
    double mma = 0, ld = 0;
    for (int i = 0; i < sms; ++i) { mma += h[i].mma_cycles; ld += h[i].ld_cycles; }
    mma /= sms; ld /= sms;
    const double ld_bytes = (double)ld_iters * ld_warps * 2 * 4096;
    printf("%-44s mma %9.0f cyc (%6.1f cyc/MMA)   ld %9.0f cyc (%6.1f B/cyc/SM)\n", name, mma,
           mma_iters ? mma / (mma_iters * 4.0) : 0.0, ld, ld > 0 ? ld_bytes / ld : 0.0);
  };
  const int MI = 2000;           // 8000 MMAs of 128x256x16
  run("mma alone N=256", MI, 0, 0, 0, 256);
  run("mma alone N=128", MI, 0, 0, 0, 128);
  for (int w : {1, 4, 8, 16}) {
    char nm[64];
    snprintf(nm, sizeof nm, "ld alone, %2d warps, serial", w);
    run(nm, 0, 4000, w, 0, 256);
    snprintf(nm, sizeof nm, "ld alone, %2d warps, 2 in flight", w);
    run(nm, 0, 4000, w, 1, 256);
  }
  // overlap: readers sized to take about as long as the MMAs do alone
  for (int w : {4, 8, 16}) {
    for (int li : {2000, 8000, 32000}) {
      char nm[64];
      snprintf(nm, sizeof nm, "mma N=256 + ld %2d warps x %5d iters", w, li / w * 4);
      run(nm, MI, li / w * 4, w, 1, 256);
    }
  }
  return 0;
}
