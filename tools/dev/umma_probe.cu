// Developer probe (B200): facts the tensor-core depthwise design rests on.
//  1. A-operand descriptors (K-major, SWIZZLE_128B) whose start address is shifted by a whole number of 128-byte rows
//     that is NOT a multiple of 8: does tcgen05.mma read rows start+r*128 with the swizzle taken from the absolute
//     address (base_offset field 0), or does it need base_offset = (start >> 7) & 7?
//  2. Cost of M128 N16 K16 / N32 / N64 / N128 / N256 tcgen05.mma issued back to back by one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I calipsync_b200/csrc tools/dev/umma_probe.cu -o build/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace casync;

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t saddr, uint32_t base_off) {
  return umma_desc_sw128(saddr) | ((uint64_t)(base_off & 7u) << 49);
}

__host__ __device__ inline int aval(int row, int col) { return ((row * 7 + col * 3) % 61) - 30; }

constexpr int kRows = 384;   // rows of the A region (128 B each)

// out[test][128][16] fp32
__global__ void __launch_bounds__(160, 1) probe_kernel(float* out, const int* shifts, int nshift, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + kRows * 128, sBAR = sB + 32768, slot = sBAR + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A: row R, chunk c (8 bf16) at R*128 + ((c ^ (R & 7)) << 4): the absolute-address SWIZZLE_128B image
  for (int i = tid; i < kRows * 64; i += blockDim.x) {
    const int R = i >> 6, col = i & 63, c = col >> 3, e = col & 7;
    *reinterpret_cast<__nv_bfloat16*>(g + R * 128 + ((c ^ (R & 7)) << 4) + e * 2) = __float2bfloat16((float)aval(R, col));
  }
  // B: 16 rows (n) x 64 k, B[n][16*grp + n] = 1 (block-diagonal identity per 16-channel group)
  for (int i = tid; i < 16 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63, c = k >> 3, e = k & 7;
    *reinterpret_cast<__nv_bfloat16*>(g + kRows * 128 + n * 128 + ((c ^ (n & 7)) << 4) + e * 2) =
        __float2bfloat16(((k & 15) == n) ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(sBAR, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  uint32_t phase = 0;
  const uint32_t idesc16 = umma_idesc_bf16(128, 16);
  int test = 0;
  for (int si = 0; si < nshift; ++si) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int grp = 0; grp < 4; grp += 3, ++test) {
        const uint32_t start = sA + shifts[si] * 128;
        if (warp == 4) {
          if (elect_one()) {
            const uint64_t ad = desc_sw128_bo(start, mode ? (start >> 7) : 0) + 2 * grp;
            const uint64_t bd = umma_desc_sw128(sB) + 2 * grp;
            umma_bf16(tmem, ad, bd, idesc16, 0);
            umma_commit(sBAR);
          }
          __syncwarp();
        }
        mbar_wait(sBAR, phase);
        phase ^= 1;
        tc_fence_after();
        if (warp < 4) {
          uint32_t r[16];
          tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
          tmem_ld_wait16(r);
          for (int j = 0; j < 16; ++j) out[((size_t)test * 128 + warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
      }
    }
  }
  // ---- timing: 72 MMAs per round (9 taps x 8 column groups), 8 rounds, one commit at the end
  const int Ns[5] = {16, 32, 64, 128, 256};
  for (int ni = 0; ni < 5; ++ni) {
    const int N = Ns[ni];
    const uint32_t idesc = umma_idesc_bf16(128, N);
    __syncthreads();
    if (warp == 4) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int round = 0; round < 8; ++round)
#pragma unroll 1
          for (int t9 = 0; t9 < 9; ++t9) {
            const uint64_t ad = umma_desc_sw128(sA + (t9 / 3 * 42 + t9 % 3) * 128);
            const uint64_t bd = umma_desc_sw128(sB);
#pragma unroll
            for (int gq = 0; gq < 8; ++gq)
              umma_bf16(tmem + (N <= 64 ? gq * N : 0), ad + 2 * (gq & 3), bd + 2 * (gq & 3), idesc, 1);
          }
        umma_commit(sBAR);
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(sBAR, phase);
      const long long t2 = clock64();
      if (lane == 0) {
        cycles[ni * 2] = t1 - t0;
        cycles[ni * 2 + 1] = t2 - t0;
      }
    } else {
      mbar_wait(sBAR, phase);
    }
    phase ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
  const std::vector<int> shifts = {0, 1, 3, 8, 17, 42, 43, 44, 84, 85, 86, 129};
  const int nshift = (int)shifts.size(), ntest = nshift * 4;
  float* dout;
  int* dsh;
  long long* dcy;
  cudaMalloc(&dout, (size_t)ntest * 128 * 16 * 4);
  cudaMalloc(&dsh, nshift * 4);
  cudaMalloc(&dcy, 80);
  cudaMemcpy(dsh, shifts.data(), nshift * 4, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + 32768 + 256 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 160, smem>>>(dout, dsh, nshift, dcy);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> h((size_t)ntest * 128 * 16);
  cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
  int test = 0;
  for (int si = 0; si < nshift; ++si)
    for (int mode = 0; mode < 2; ++mode)
      for (int grp = 0; grp < 4; grp += 3, ++test) {
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n) {
            const float want = (float)aval(m + shifts[si], 16 * grp + n);
            if (h[((size_t)test * 128 + m) * 16 + n] != want) {
              if (first < 0) first = m * 16 + n;
              ++bad;
            }
          }
        printf("shift %3d base_offset=%s k-group %d: %4d / 2048 mismatches%s", shifts[si], mode ? "(start>>7)&7" : "0", grp,
               bad, bad ? "" : "  OK");
        if (bad) {
          const int m = first / 16, n = first % 16;
          printf("  first at row %d col %d: got %g want %d", m, n, h[((size_t)test * 128 + m) * 16 + n],
                 aval(m + shifts[si], 16 * grp + n));
        }
        printf("\n");
      }
  long long cy[10];
  cudaMemcpy(cy, dcy, 80, cudaMemcpyDeviceToHost);
  const int Ns[5] = {16, 32, 64, 128, 256};
  for (int i = 0; i < 5; ++i)
    printf("576 x tcgen05.mma M128 N%-3d K16: issue %lld cycles (%.1f / mma), complete %lld cycles (%.1f / mma)\n", Ns[i],
           cy[2 * i], cy[2 * i] / 576.0, cy[2 * i + 1], cy[2 * i + 1] / 576.0);
  return 0;
}
