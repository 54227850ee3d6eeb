// Strip-streaming fused InvertedResidual with the depthwise 3x3 ON THE TENSOR CORES, for the blocks with 64 hidden
// channels (up4.1, up3.1, audio conv1).  Same geometry as strip_ir.cu (strips of SW output columns, a contiguous range of
// the global padded-row list per CTA, hidden positions in 128-row tiles), but the depthwise conv is 9 taps x 4 channel
// groups of
//     DW[tile][:, 16g..16g+15] += HID[rows shifted by dy*WW+dx][:, 16g..16g+15] . diag(w[tap, 16g..16g+15])
// i.e. tcgen05.mma M128 N16 K16 whose A operand is the hidden tile read at a ROW OFFSET (a K-major SWIZZLE_128B tile
// may start at any 128-byte row: the swizzle is a function of the absolute shared-memory address, tools/dev/
// umma_probe.cu) and whose B operand is a 16x16 diagonal block of taps.  15/16 of those MACs are zeros and an N=16 MMA
// is bound by re-reading its 4 KB A operand (39 cycles, measured), so this costs 36 x 39 = 1400 cycles per 128-position
// tile -- but it costs NO issue slots: on the CUDA cores the same depthwise pass is a 520-instruction serial chain per
// thread per tile (2900 cycles), the longest role of strip_ir.cu for these blocks.
//
//   producers (4 warps)  A1[t]  <- global (cp.async, zero fill outside the image)
//   issuer 0             D1[t]  =  A1[t] . W1^T                 M128 N64 K16 x CIN/16        -> TMEM
//   drain 1 (8 warps)    HID[t] =  leaky(D1 + b1) bf16, 0 outside the image -> ring of 3 SWIZZLE_128B tiles (+ a copy of
//                                  the first rows of slot 0 behind slot 2, so that a shifted view never wraps)
//   issuer 1             DW[t]  =  sum over 9 taps, 4 groups (see above); needs HID[t] and HID[t+1]   -> TMEM
//   drain 2 (8 warps)    A2[t]  =  leaky(DW + bd) bf16 -> SWIZZLE_128B tile
//   issuer 2             D2[t]  =  A2[t] . W2^T                 M128 N=COUT K16 x 4          -> TMEM
//   epilogue (4 warps)   out    =  leaky(D2 + b2) (+ x) for the positions that are outputs (hx < SW, row is a top row)
#include "strip_ir.cuh"

#include <cuda_bf16.h>

#include <cstdlib>

namespace casync {

namespace {

#ifndef STRIP_TC_ORDER
#define STRIP_TC_ORDER 1
#endif
#if STRIP_TC_ORDER
// the warp arbiter favours high warp ids: the three MMA issuers (a few instructions per tile, but every other role waits
// for them) sit at the top
constexpr int kProdWarp0 = 0, kEpiWarp0 = 4, kDrain1Warp0 = 8, kDrain2Warp0 = 16, kIssWarp0 = 24;
#else
constexpr int kIssWarp0 = 0, kEpiWarp0 = 4, kDrain1Warp0 = 8, kDrain2Warp0 = 16, kProdWarp0 = 24;
#endif
constexpr int kThreads = 28 * 32;
constexpr int kTileB = 16384;
constexpr int kSleepNs = 64;
#ifndef STRIP_EXP
#define STRIP_EXP 0   // developer timing experiments (wrong results): 1 one tap only, 8 no epilogue stores / residual loads
#endif

template <int CIN_, int COUT_, int W_, int SW_, bool RES_>
struct TCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, W = W_, SW = SW_;
  static constexpr bool RES = RES_;
  static constexpr int CH = 2 * CIN, WW = SW + 2, S = W / SW, HP = W + 2;
  static constexpr int OV = (2 * WW + 2 + 7) & ~7;          // rows a shifted view reads beyond its tile
  static constexpr int oA1 = 0;
  static constexpr int oHID = oA1 + 2 * kTileB;
  static constexpr int oA2 = oHID + 3 * kTileB + OV * 128;
  static constexpr int oW1 = oA2 + 2 * kTileB;
  static constexpr int oW2 = oW1 + CH * 128;
  static constexpr int oWD = oW2 + COUT * 128;              // 9 taps x [16 rows x 128 B]: diagonal blocks, group g at +32g
  static constexpr int oMETA = oWD + 9 * 2048;
  static constexpr int oBAR = oMETA + 4 * 128;
  static constexpr int kSmem = oBAR + 512 + 1024;
  static constexpr uint32_t kWeightBytes = CH * 128 + COUT * 128;
  static_assert(CH == 64, "one 64-channel chunk of hidden channels");
  static_assert(W % SW == 0 && 2 * WW + 2 <= 128, "a shifted view may reach into the next tile only");
  static_assert(COUT <= 128 && 2 * 64 + 2 * 64 + 2 * COUT <= 512, "TMEM");
  static_assert(kSmem <= 232448, "shared memory overflow");
};

enum Bar : int {
  B_W = 0, B_A1FULL = 1, B_A1FREE = 3, B_D1FULL = 5, B_D1FREE = 7, B_HIDFULL = 9, B_HIDFREE = 12, B_DWFULL = 15,
  B_DWFREE = 17, B_A2FULL = 19, B_A2FREE = 21, B_D2FULL = 23, B_D2FREE = 25, B_COUNT = 27
};

struct SmemView {
  uint8_t* g;
  uint32_t base;
  template <class T>
  __device__ __forceinline__ T& at(uint32_t addr) const { return *reinterpret_cast<T*>(g + (addr - base)); }
};

// leaky(acc + bias) for 8 consecutive accumulator columns -> 4 packed bf16x2 (bias: constant-bank operands)
__device__ __forceinline__ void bias_leaky8(const uint32_t* acc, const float* __restrict__ bias, const __nv_bfloat162 kslope,
                                            uint32_t* o) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[2 * j]) + bias[2 * j],
                                             __uint_as_float(acc[2 * j + 1]) + bias[2 * j + 1]);
    v = __hmax2(v, __hmul2(v, kslope));
    o[j] = *reinterpret_cast<uint32_t*>(&v);
  }
}

template <class C>
__global__ void __launch_bounds__(kThreads, 1) strip_tc_kernel(const __grid_constant__ StripArgs p) {
  constexpr int CIN = C::CIN, COUT = C::COUT, W = C::W, SW = C::SW, CH = C::CH, WW = C::WW, S = C::S, HP = C::HP,
                H = C::W, OV = C::OV;
  constexpr bool RES = C::RES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const SmemView sm{smem_raw + (base - smem_u32(smem_raw)), base};
  const uint32_t sA1 = base + C::oA1, sHID = base + C::oHID, sA2 = base + C::oA2, sW1 = base + C::oW1,
                 sW2 = base + C::oW2, sWD = base + C::oWD, sMETA = base + C::oMETA, sBAR = base + C::oBAR;
  auto bar = [&](int i) { return sBAR + 8u * i; };
  const uint32_t tmem_slot = sBAR + 8u * B_COUNT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();

  if (tid == 0) {
    mbar_init(bar(B_W), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_A1FULL + i), 128);
      mbar_init(bar(B_A1FREE + i), 1);
      mbar_init(bar(B_D1FULL + i), 1);
      mbar_init(bar(B_D1FREE + i), 256);
      mbar_init(bar(B_DWFULL + i), 1);
      mbar_init(bar(B_DWFREE + i), 256);
      mbar_init(bar(B_A2FULL + i), 256);
      mbar_init(bar(B_A2FREE + i), 1);
      mbar_init(bar(B_D2FULL + i), 1);
      mbar_init(bar(B_D2FREE + i), 128);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar(B_HIDFULL + i), 256);
      mbar_init(bar(B_HIDFREE + i), 1);
    }
    fence_mbar_init();
    mbar_arrive_expect_tx(bar(B_W), C::kWeightBytes);
    bulk_g2s(sW1, p.W1, CH * 128, bar(B_W));
    bulk_g2s(sW2, p.W2, COUT * 128, bar(B_W));
  }
  if (warp == kIssWarp0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // diagonal tap blocks: tap t9 -> [16 rows n][64 k] bf16, SWIZZLE_128B image; element (n, 16g + n) = w[t9][16g + n]
  for (int i = tid; i < 9 * 2048 / 16; i += kThreads) sm.at<uint4>(sWD + i * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < 9 * 64) {
    const int t9 = tid >> 6, ch = tid & 63, g = ch >> 4, n = ch & 15;
    // wdp: bf16 [CH/8][10][8]
    const __nv_bfloat16 wv = reinterpret_cast<const __nv_bfloat16*>(p.wdp)[((ch >> 3) * 10 + t9) * 8 + (ch & 7)];
    const int k = 16 * g + n;   // column of row n
    sm.at<__nv_bfloat16>(sWD + t9 * 2048 + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = wv;
  }
  fence_proxy_async();   // generic-proxy writes of WD -> async proxy (tcgen05.mma reads)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tD1 = tmem, tDW = tmem + 128, tD2 = tmem + 256;

  const long long NG = (long long)p.batch * S * HP;
  const int G0 = (int)(NG * blockIdx.x / gridDim.x), G1 = (int)(NG * (blockIdx.x + 1) / gridDim.x);
  const int ntop = G1 - G0, nrows = ntop + 2;
  const int NT = (nrows * WW + 127) >> 7;
  const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);

  if (warp >= kIssWarp0 && warp < kIssWarp0 + 4) {
    if (warp == kIssWarp0) {
      // ----- first 1x1 conv
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, CH);
      mbar_wait_sleep<kSleepNs>(bar(B_W), 0);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1;
        mbar_wait_sleep<kSleepNs>(bar(B_A1FULL + s), ph);
        mbar_wait_sleep<kSleepNs>(bar(B_D1FREE + s), ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(sA1 + s * kTileB), bd = umma_desc_sw128(sW1);
#pragma unroll
          for (int ks = 0; ks < CIN / 16; ++ks) umma_bf16(tD1 + s * 64, ad + 2 * ks, bd + 2 * ks, idesc1, ks != 0);
          umma_commit(bar(B_D1FULL + s));
          umma_commit(bar(B_A1FREE + s));
        }
        __syncwarp();
      }
    } else if (warp == kIssWarp0 + 1) {
      // ----- depthwise 3x3 as 36 shifted N16 MMAs per tile
      constexpr uint32_t idesc16 = umma_idesc_bf16(128, 16);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1, hs = t % 3;
        mbar_wait_sleep<kSleepNs>(bar(B_HIDFULL + hs), (t / 3) & 1);
        if (t + 1 < NT) mbar_wait_sleep<kSleepNs>(bar(B_HIDFULL + (t + 1) % 3), ((t + 1) / 3) & 1);
        mbar_wait_sleep<kSleepNs>(bar(B_DWFREE + s), ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a0 = umma_desc_sw128(sHID + hs * kTileB);
          const uint64_t b0 = umma_desc_sw128(sWD);
#pragma unroll
          for (int t9 = 0; t9 < ((STRIP_EXP & 1) ? 1 : 9); ++t9) {
            const uint64_t ad = a0 + (uint64_t)(((t9 / 3) * WW + (t9 % 3)) * 128 >> 4);   // start address += shift rows
            const uint64_t bd = b0 + (uint64_t)(t9 * 2048 >> 4);
#pragma unroll
            for (int g = 0; g < 4; ++g) umma_bf16(tDW + s * 64 + g * 16, ad + 2 * g, bd + 2 * g, idesc16, t9 != 0);
          }
          umma_commit(bar(B_DWFULL + s));
          umma_commit(bar(B_HIDFREE + hs));
        }
        __syncwarp();
      }
    } else if (warp == kIssWarp0 + 2) {
      // ----- second 1x1 conv
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, COUT);
      mbar_wait_sleep<kSleepNs>(bar(B_W), 0);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1;
        mbar_wait_sleep<kSleepNs>(bar(B_A2FULL + s), ph);
        mbar_wait_sleep<kSleepNs>(bar(B_D2FREE + s), ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(sA2 + s * kTileB), bd = umma_desc_sw128(sW2);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tD2 + s * COUT, ad + 2 * ks, bd + 2 * ks, idesc2, ks != 0);
          umma_commit(bar(B_A2FREE + s));
          umma_commit(bar(B_D2FULL + s));
        }
        __syncwarp();
      }
    }
  } else if (warp >= kProdWarp0 && warp < kProdWarp0 + 4) {
    // =========================================== producers: A1[t] <- global ====================================
    pdl_wait();
    const int r = tid - kProdWarp0 * 32;
    const uint32_t r7 = r & 7;
    for (int t = 0; t <= NT; ++t) {
      if (t < NT) {
        const int s = t & 1;
        if (t >= 2) mbar_wait_sleep<kSleepNs>(bar(B_A1FREE + s), ((t >> 1) & 1) ^ 1);
        const int q = t * 128 + r;
        const int jrow = q / WW, hx = q - jrow * WW;
        const int G = G0 + jrow;
        const int bs = G / HP, hy = G - bs * HP;
        const int b = bs / S, st = bs - b * S;
        const int y = hy - 1, x = st * SW + hx - 1;
        const bool inside = jrow < nrows && bs < p.batch * S && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        sm.at<uint8_t>(sMETA + (t & 3) * 128 + r) = inside ? 1 : 0;
        const uint32_t a1 = sA1 + s * kTileB + r * 128;
        const size_t pix = inside ? ((size_t)b * H + y) * W + x : 0;
        const __nv_bfloat16* src = p.in + pix * CIN;
#pragma unroll
        for (int c = 0; c < CIN / 8; ++c) cp_async16(a1 + ((c ^ r7) << 4), src + c * 8, inside);
        cp_async_commit();
      }
      if (t >= 1) {   // tile t-1 has landed once at most the newest group is still in flight
        if (t < NT) cp_async_wait<1>(); else cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(bar(B_A1FULL + ((t - 1) & 1)));
      }
    }
  } else if (warp >= kDrain1Warp0 && warp < kDrain1Warp0 + 8) {
    // =========================================== drain 1: D1 -> HID ring =======================================
    const int dwp = warp - kDrain1Warp0, lg = dwp & 3, hw = dwp >> 2;   // TMEM lane quarter, column half (32 columns)
    const int row = lg * 32 + lane;
    const uint32_t r7 = row & 7;
    for (int t = 0; t < NT; ++t) {
      const int b = t & 1, hs = t % 3;
      mbar_wait_sleep<kSleepNs>(bar(B_D1FULL + b), (t >> 1) & 1);
      if (t >= 3) mbar_wait_sleep<kSleepNs>(bar(B_HIDFREE + hs), ((t / 3) & 1) ^ 1);
      const bool inside = sm.at<uint8_t>(sMETA + (t & 3) * 128 + row) != 0;
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld32(tD1 + b * 64 + hw * 32 + ((uint32_t)(lg * 32) << 16), acc);
      tmem_ld_wait32(acc);
      tc_fence_before();
      mbar_arrive(bar(B_D1FREE + b));
      uint32_t o[4][4];
      if (hw == 0) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) bias_leaky8(acc + g8 * 8, p.b1 + g8 * 8, kslope, o[g8]);
      } else {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) bias_leaky8(acc + g8 * 8, p.b1 + 32 + g8 * 8, kslope, o[g8]);
      }
      const uint32_t hid = sHID + hs * kTileB + row * 128;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const uint32_t j = (uint32_t)(hw * 4 + g8);
        const uint4 v = inside ? make_uint4(o[g8][0], o[g8][1], o[g8][2], o[g8][3]) : make_uint4(0, 0, 0, 0);
        sm.at<uint4>(hid + ((j ^ r7) << 4)) = v;
        if (hs == 0 && row < OV) sm.at<uint4>(hid + 3 * kTileB + ((j ^ r7) << 4)) = v;   // copy behind slot 2
      }
      fence_proxy_async();   // HID is an MMA operand
      mbar_arrive(bar(B_HIDFULL + hs));
    }
  } else if (warp >= kDrain2Warp0 && warp < kDrain2Warp0 + 8) {
    // =========================================== drain 2: DW -> A2 ==============================================
    const int dwp = warp - kDrain2Warp0, lg = dwp & 3, hw = dwp >> 2;
    const int row = lg * 32 + lane;
    const uint32_t r7 = row & 7;
    // folded-BN bias of the depthwise conv for this thread's 32 channels: packed bf16 pairs as stored in wdp (row 9 of
    // every 8-channel entry); added after the fp32 accumulator is rounded to bf16 (the taps themselves are bf16)
    __nv_bfloat162 bd2[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 b8 = __ldg(reinterpret_cast<const uint4*>(p.wdp + (size_t)(hw * 4 + i) * 160 + 9 * 16));
      const uint32_t* pb = &b8.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) bd2[i * 4 + j] = *reinterpret_cast<const __nv_bfloat162*>(&pb[j]);
    }
    for (int t = 0; t < NT; ++t) {
      const int b = t & 1;
      mbar_wait_sleep<kSleepNs>(bar(B_DWFULL + b), (t >> 1) & 1);
      if (t >= 2) mbar_wait_sleep<kSleepNs>(bar(B_A2FREE + b), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld32(tDW + b * 64 + hw * 32 + ((uint32_t)(lg * 32) << 16), acc);
      tmem_ld_wait32(acc);
      tc_fence_before();
      mbar_arrive(bar(B_DWFREE + b));
      const uint32_t a2 = sA2 + b * kTileB + row * 128;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[g8 * 8 + 2 * j]), __uint_as_float(acc[g8 * 8 + 2 * j + 1]));
          v = __hadd2(v, bd2[g8 * 4 + j]);
          v = __hmax2(v, __hmul2(v, kslope));
          o[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        const uint32_t j = (uint32_t)(hw * 4 + g8);
        sm.at<uint4>(a2 + ((j ^ r7) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async();
      mbar_arrive(bar(B_A2FULL + b));
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // =========================================== epilogue: D2 -> global ========================================
    pdl_wait();
    const int row = (warp - kEpiWarp0) * 32 + lane;
    const bool fuse_outc = COUT == 32 && p.final_out != nullptr;   // last decoder block: the output head runs right here
    for (int t = 0; t < NT; ++t) {
      const int s = t & 1;
      const int q = t * 128 + row;
      const int jrow = q / WW, hx = q - jrow * WW;
      const int G = G0 + jrow;
      const int bs = G / HP, hy = G - bs * HP;
      const int b = bs / S, st = bs - b * S;
      const bool valid = jrow < ntop && hy < H && hx < SW && bs < p.batch * S;
      const size_t pix = valid ? ((size_t)b * H + hy) * W + st * SW + hx : 0;
      uint4 rr[RES ? COUT / 8 : 1];
      if constexpr (RES) {
#pragma unroll
        for (int i = 0; i < COUT / 8; ++i)
          rr[i] = (valid && !(STRIP_EXP & 8)) ? __ldg(reinterpret_cast<const uint4*>(p.in + pix * CIN + i * 8))
                                              : make_uint4(0, 0, 0, 0);
      }
      mbar_wait_sleep<kSleepNs>(bar(B_D2FULL + s), (t >> 1) & 1);
      tc_fence_after();
      float a0 = p.bo[0], a1 = p.bo[1], a2 = p.bo[2];   // OutConv accumulators (used when fuse_outc)
#pragma unroll
      for (int cc = 0; cc < COUT; cc += 32) {
        uint32_t acc[32];
        tmem_ld32(tD2 + s * COUT + cc + ((uint32_t)((warp & 3) * 32) << 16), acc);
        tmem_ld_wait32(acc);
        if (cc + 32 >= COUT) {
          tc_fence_before();
          mbar_arrive(bar(B_D2FREE + s));
        }
        if (valid && !((STRIP_EXP & 8) && acc[0] != 0x12345u)) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float vv[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              vv[jj] = __uint_as_float(acc[g8 * 8 + jj]) + p.b2[cc + g8 * 8 + jj];
              vv[jj] = fmaxf(vv[jj], kLeaky * vv[jj]);
            }
            if constexpr (RES) {
              const uint32_t* pr = &rr[(cc >> 3) + g8].x;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                vv[2 * jj] += bf16_lo(pr[jj]);
                vv[2 * jj + 1] += bf16_hi(pr[jj]);
              }
            }
            const uint4 o = make_uint4(pack_bf16(vv[0], vv[1]), pack_bf16(vv[2], vv[3]), pack_bf16(vv[4], vv[5]),
                                       pack_bf16(vv[6], vv[7]));
            if constexpr (COUT == 32) {
              if (fuse_outc) {
                // OutConv on the ROUNDED bf16 values in the order of outc_kernel: bit-identical to the two-launch path
                const uint32_t* pv = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float lo = bf16_lo(pv[j]), hi = bf16_hi(pv[j]);
                  const int c = g8 * 8 + 2 * j;
                  a0 = fmaf(p.wo[c + 1], hi, fmaf(p.wo[c], lo, a0));
                  a1 = fmaf(p.wo[32 + c + 1], hi, fmaf(p.wo[32 + c], lo, a1));
                  a2 = fmaf(p.wo[64 + c + 1], hi, fmaf(p.wo[64 + c], lo, a2));
                }
                continue;
              }
            }
            *reinterpret_cast<uint4*>(p.out + pix * COUT + cc + g8 * 8) = o;
          }
        }
      }
      if constexpr (COUT == 32) {
        if (fuse_outc && valid) {
          const float s0 = 1.f / (1.f + __expf(-a0)), s1 = 1.f / (1.f + __expf(-a1)), s2 = 1.f / (1.f + __expf(-a2));
          if (p.final_u8) {   // floor(p * 255) like `np.array(pred * 255, dtype=np.uint8)` (infer_api.py:265-266)
            uint8_t* o8 = reinterpret_cast<uint8_t*>(p.final_out) + pix * 3;
            o8[0] = (uint8_t)(s0 * 255.f);
            o8[1] = (uint8_t)(s1 * 255.f);
            o8[2] = (uint8_t)(s2 * 255.f);
          } else {
            float* of = reinterpret_cast<float*>(p.final_out) + (size_t)b * 76800 + (size_t)hy * W + st * SW + hx;
            of[0] = s0;
            of[25600] = s1;
            of[51200] = s2;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssWarp0) tmem_dealloc(tmem, 512);
}

template <class C>
int launch_t(const StripArgs& a, int num_sms, cudaStream_t st) {
  static unsigned long long attr_devs = 0;
  auto kfn = strip_tc_kernel<C>;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_devs >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_devs |= 1ull << (dev & 63);
  }
  const long long NG = (long long)a.batch * C::S * C::HP;
  long long grid = NG / 6;
  if (grid < 1) grid = 1;
  if (grid > num_sms) grid = num_sms;
  return (int)launch_pdl(kfn, dim3((unsigned)grid), dim3(kThreads), C::kSmem, st, a);
}

}  // namespace

#define STRIP_TC_CASES(X)                                  \
  X(32, 32, 160, 40, true)   /* up4.1 */                  \
  X(32, 32, 80, 40, true)    /* up3.1 */                  \
  X(32, 64, 32, 32, false)   /* audio conv1 */

bool strip_tc_supported(int cin, int cout, int W, int stride, bool upcat, bool res) {
  if (stride != 1 || upcat) return false;
#define X(CIN_, COUT_, W_, SW_, R_) \
  if (cin == CIN_ && cout == COUT_ && W == W_ && res == R_) return true;
  STRIP_TC_CASES(X)
#undef X
  return false;
}

int launch_strip_tc(const StripArgs& a, int cin, int cout, int W, int stride, bool upcat, bool res, int num_sms,
                    cudaStream_t st) {
  if (stride != 1 || upcat) return -1;
#define X(CIN_, COUT_, W_, SW_, R_)                                       \
  if (cin == CIN_ && cout == COUT_ && W == W_ && res == R_)               \
    return launch_t<TCfg<CIN_, COUT_, W_, SW_, R_>>(a, num_sms, st);
  STRIP_TC_CASES(X)
#undef X
  return -1;
}

}  // namespace casync
