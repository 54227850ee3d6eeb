// Developer microbenchmark: L2 -> SM read bandwidth with (a) LDG.128 and (b) 1-D bulk-async copies (UBLKCP),
// over a working set that stays L2 resident; plus (c) bulk copies multicast to a 2-/4-CTA cluster.
// Numbers size the tile shapes of gemm_tc / fused kernels (bytes per MMA cycle an SM can be fed from L2).
#include <cstdio>
#include <cooperative_groups.h>
#include "../../calipsync_b200/csrc/common.cuh"
using namespace casync;
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ src, size_t n16, int iters, uint32_t* sink) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = i + u * stride < n16 ? __ldcg(src + i + u * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  if (acc == 0x9e3779b9u) *sink = acc;
}

// each CTA streams `chunk`-byte pieces into a ring of shared-memory stages with bulk copies
template <int CLUSTER>
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ src, size_t bytes, int chunk, int iters,
                                                   uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int S = 4;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + S * chunk;
  uint32_t rank = 0;
  if (CLUSTER > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    fence_mbar_init();
  }
  if (CLUSTER > 1) cg::this_cluster().sync(); else __syncthreads();
  const size_t nchunks = bytes / chunk;
  const size_t cl = blockIdx.x / CLUSTER, ncl = gridDim.x / CLUSTER;
  if (threadIdx.x == 0) {
    size_t j = 0;
    for (int it = 0; it < iters; ++it)
      for (size_t c = cl; c < nchunks; c += ncl, ++j) {
        const int s = j % S;
        if (j >= S) mbar_wait(bars + 8 * s, ((j / S) - 1) & 1);
        mbar_arrive_expect_tx(bars + 8 * s, chunk);
        if (CLUSTER == 1) {
          bulk_g2s(base + s * chunk, src + c * chunk, chunk, bars + 8 * s);
        } else {
          // each CTA of the cluster fetches 1/CLUSTER of the chunk and multicasts it to all CTAs of the cluster
          const uint32_t part = chunk / CLUSTER;
          const uint16_t mask = (1u << CLUSTER) - 1;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
              ::"r"(base + s * chunk + rank * part), "l"(src + c * chunk + rank * part), "r"(part), "r"(bars + 8 * s),
              "h"(mask) : "memory");
        }
      }
    for (size_t q = (j > S ? j - S : 0); q < j; ++q) mbar_wait(bars + 8 * (q % S), (q / S) & 1);
  }
  if (CLUSTER > 1) cg::this_cluster().sync(); else __syncthreads();
  if (threadIdx.x == 1 && sink == nullptr) *sink = 0;
}

template <int CLUSTER>
void run_bulk(const uint8_t* d, size_t bytes, int chunk) {
  cudaFuncSetAttribute(bulk_kernel<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * chunk + 2048);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 / CLUSTER * CLUSTER);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 4 * chunk + 2048;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int iters = 20;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  uint32_t* sink = nullptr;
  cudaMalloc(&sink, 4);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, bulk_kernel<CLUSTER>, d, bytes, chunk, iters, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 1)
      printf("bulk  cluster=%d chunk=%6d B  working set %.0f MB: %.0f GB/s delivered to smem (L2 reads %.0f GB/s) (%s)\n",
             CLUSTER, chunk, bytes / 1e6, (double)bytes * iters * CLUSTER / (ms * 1e-3) / 1e9,
             (double)bytes * iters / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
  }
}

int main() {
  for (size_t mb : {16, 48, 96}) {
    const size_t bytes = mb << 20;
    uint8_t* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    uint32_t* sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20;
    for (int grid : {148, 296, 592}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ldg_kernel<<<grid, 512>>>(reinterpret_cast<const uint4*>(d), bytes / 16, iters, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 1)
          printf("ldg   grid=%4d  working set %zu MB: %.0f GB/s (%s)\n", grid, mb, (double)bytes * iters / (ms * 1e-3) / 1e9,
                 cudaGetErrorString(e)), fflush(stdout);
      }
    }
    for (int chunk : {16384, 49152}) {
      run_bulk<1>(d, bytes, chunk);
      fflush(stdout);
    }
    cudaFree(d);
  }
  return 0;
}
