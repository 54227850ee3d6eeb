// Developer microbenchmark: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32) per SM vs #warps.
#include <cstdio>
#include "../../calipsync_b200/csrc/common.cuh"
using namespace casync;
__global__ void k(int iters, unsigned long long* out, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld32(base + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 8192 * 8);
  for (int nw : {1, 4, 8, 16}) {
    const int iters = 2000;
    k<<<148, 512>>>(iters, d, nw);
    cudaDeviceSynchronize();
    k<<<148, 512>>>(iters, d, nw);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)nw * iters * 32 * 32 * 4;
    printf("warps=%2d  cycles=%llu  TMEM read %.1f B/cycle/SM (%s)\n", nw, h, bytes / h, cudaGetErrorString(e));
  }
  return 0;
}
