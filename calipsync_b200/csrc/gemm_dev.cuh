// Device helpers of the GEMM kernel (gemm_tc.cu): bilinear taps of the decoder's upsample + concat producer, 2-D TMA tensor copies.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace casync {

struct RowCoord {  // UPCAT per-thread bilinear taps
  const __nv_bfloat16 *p00, *p01, *p10, *p11;
  float wy0, wy1, wx0, wx1;
};

__device__ __forceinline__ uint4 lerp8(const uint4& a, const uint4& b, const uint4& c, const uint4& d,
                                       const RowCoord& rc) {
  const uint32_t* pa = &a.x;
  const uint32_t* pb = &b.x;
  const uint32_t* pc = &c.x;
  const uint32_t* pd = &d.x;
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // same association as PyTorch's upsample_bilinear2d: wy0*(wx0*p00 + wx1*p01) + wy1*(wx0*p10 + wx1*p11)
    float lo = rc.wy0 * (rc.wx0 * bf16_lo(pa[i]) + rc.wx1 * bf16_lo(pb[i])) +
               rc.wy1 * (rc.wx0 * bf16_lo(pc[i]) + rc.wx1 * bf16_lo(pd[i]));
    float hi = rc.wy0 * (rc.wx0 * bf16_hi(pa[i]) + rc.wx1 * bf16_hi(pb[i])) +
               rc.wy1 * (rc.wx0 * bf16_hi(pc[i]) + rc.wx1 * bf16_hi(pd[i]));
    po[i] = pack_bf16(lo, hi);
  }
  return o;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// TMA tensor store shared -> global (bulk async-group of the issuing thread); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
               "r"(src)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// every bulk store of this thread has finished READING shared memory (the slab may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace casync
