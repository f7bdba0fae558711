// Microbenchmark / semantics probe for a CTA-pair (cta_group::2) GEMM in a cluster of several
// pairs with the A tile multicast across pairs.  Questions it answers on the
// hardware:
//   1. how shared::cta addresses relate to shared::cluster addresses (is the CTA rank encoded in
//      the address, so that clearing bit 24 of an odd CTA's barrier address names its pair leader's?)
//   2. does cp.async.bulk.tensor ... .cta_group::2 ... .multicast::cluster deliver the tile to every
//      CTA of the mask AND complete_tx on the pair LEADER's mbarrier of each destination CTA when
//      the barrier operand has the peer bit cleared?
//   3. how many 6-CTA clusters can be co-resident (cudaOccupancyMaxActiveClusters)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cluster_multicast cluster_multicast.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int CLUSTER>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int* result) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tile = smem;                                  // 128 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  const uint32_t rank = ctarank();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    printf("rank %u: shared::cta addr of bar = 0x%08x; mapa(.,0..%d) =", rank, smem_u32(bar), CLUSTER - 1);
    for (int r = 0; r < CLUSTER; ++r) printf(" 0x%08x", mapa(smem_u32(bar), r));
    printf("\n");
    for (int i = 0; i < 16384 / 4; ++i) reinterpret_cast<uint32_t*>(tile)[i] = 0xDEADBEEFu;
  }
  __syncthreads();
  cluster_sync();
  // Every pair leader (even rank) expects 2 x 16 KB: the tile multicast into its own and its peer's
  // shared memory.  CTA `c` (c = 0 or 1: the rank-in-pair) of pair 0 loads the tile for all CTAs of
  // the same rank-in-pair.
  if (threadIdx.x == 0) {
    if ((rank & 1) == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32768) : "memory");
    }
    if (rank < 2) {
      uint16_t mask = 0;
      for (int r = rank; r < CLUSTER; r += 2) mask |= (uint16_t)(1u << r);
      const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;      // peer bit cleared: the pair leader's
      asm volatile(
          "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
          " [%0], [%1, {%3, %4}], [%2], %5;"
          ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar_addr), "r"(0), "r"((int)rank * 128), "h"(mask)
          : "memory");
    }
    if ((rank & 1) == 0) {
      // bounded wait on the leader's barrier
      uint32_t ok = 0;
      long long t0 = clock64();
      while (!ok && clock64() - t0 < 2000000000ll) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
      }
      result[blockIdx.x * 4 + 0] = (int)ok;
    }
  }
  __syncthreads();
  cluster_sync();       // leaders have seen the bytes of both CTAs of their pair (or timed out)
  if (threadIdx.x == 0) {
    const uint32_t* t = reinterpret_cast<const uint32_t*>(tile);
    // with 128B swizzle row 0 / chunk 0 stays at offset 0: element (row, 0..1) as a bf16 pair
    result[blockIdx.x * 4 + 1] = (int)t[0];
    result[blockIdx.x * 4 + 2] = (int)t[(127 * 128) / 4];        // row 127, first word (chunk 0 ^ 7 -> check loosely)
    result[blockIdx.x * 4 + 3] = (int)rank;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CLUSTER>
static void run(const CUtensorMap& tm, int* d_res) {
  auto kern = probe_kernel<CLUSTER>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64 + 1024));
  if (CLUSTER > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CLUSTER);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = 16384 + 64 + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n_clusters = -1;
  cudaLaunchConfig_t occ = cfg;
  occ.gridDim = dim3(CLUSTER * 64);
  occ.dynamicSmemBytes = 200 * 1024;        // what the real GEMM uses: one CTA per SM
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &occ);
  printf("cluster size %d: cudaOccupancyMaxActiveClusters (200 KB smem/CTA) = %d (%s) -> %d SMs\n", CLUSTER,
         n_clusters, cudaGetErrorString(e), n_clusters * CLUSTER);
  CK(cudaMemset(d_res, 0xff, 64 * 4 * sizeof(int)));
  CK(cudaLaunchKernelEx(&cfg, kern, tm, d_res));
  e = cudaDeviceSynchronize();
  printf("cluster size %d: kernel -> %s\n", CLUSTER, cudaGetErrorString(e));
  int h[64 * 4];
  CK(cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost));
  for (int b = 0; b < CLUSTER; ++b)
    printf("  cta %d (rank %d): leader wait ok = %d, tile[0] = 0x%08x (expect row %d*128 pattern 0x%04x%04x)\n", b, h[b * 4 + 3],
           h[b * 4 + 0], (unsigned)h[b * 4 + 1], b & 1, ((b & 1) * 128 * 64 + 1) & 0xffff, ((b & 1) * 128 * 64) & 0xffff);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  // A: 256 rows x 64 u16 (128 B per row); element (r, c) = r*64 + c
  const int rows = 256, cols = 64;
  uint16_t* h = (uint16_t*)malloc(rows * cols * 2);
  for (int i = 0; i < rows * cols; ++i) h[i] = (uint16_t)i;
  void* d;
  CK(cudaMalloc(&d, rows * cols * 2));
  CK(cudaMemcpy(d, h, rows * cols * 2, cudaMemcpyHostToDevice));
  int* d_res;
  CK(cudaMalloc(&d_res, 64 * 4 * sizeof(int)));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  run<2>(tm, d_res);
  run<4>(tm, d_res);
  run<6>(tm, d_res);
  run<8>(tm, d_res);
  return 0;
}
