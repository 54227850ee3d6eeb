// Strip-streaming fused InvertedResidual with the depthwise 3x3 ON THE TENSOR CORES, for the blocks with 64 hidden
// channels (up4.1 -- with the network's output head in its epilogue --, up3.1, audio conv1) and, as the INC instantiation
// (16 padded hidden channels, fp32 NCHW input; see TCfg), for the input block.  Same geometry as strip_ir.cu (strips of SW output columns, a contiguous range of
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
constexpr int kTileB = 16384;
#ifndef STRIP_TC_SLEEP
#define STRIP_TC_SLEEP 64
#endif
constexpr int kSleepNs = STRIP_TC_SLEEP;
#ifndef STRIP_DBG
#define STRIP_DBG 0   // 1: per-role cycle counters (CASYNC_PHASE_DBG=<ir index>, 0 = the input block)
#endif
#ifndef STRIP_TC_INC_LDG
#define STRIP_TC_INC_LDG 1   // input block: 1 = LDG into registers one tile ahead, 0 = 4-byte cp.async staging ring
#endif
#ifndef STRIP_TC_ONE_COMMIT
// 1: one tcgen05.commit per MMA group.  "operand slot free" and "accumulator full" are the same event (the group has
// completed), so the roles that refill A1 / HID / A2 wait on D1FULL / DWFULL / D2FULL with the parity of the tile that
// used the slot; those barriers cannot run ahead of such a waiter because the next group on the same barrier needs the
// waiter's own output first.
#define STRIP_TC_ONE_COMMIT 1
#endif
#ifndef STRIP_EXP
#define STRIP_EXP 0   // developer timing experiments (wrong results): 1 one tap only, 8 no epilogue stores / residual loads
#endif

template <int CIN_, int COUT_, int W_, int SW_, bool RES_, int CH_ = 2 * CIN_, bool INC_ = false, bool EG2_ = false>
struct TCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, W = W_, SW = SW_;
  static constexpr bool RES = RES_, INC = INC_;
  static constexpr bool EG2 = EG2_;                         // a second epilogue group in four extra warps (28..31)
  static constexpr int kThreads = (EG2 ? 32 : 28) * 32;
  // INC: the 6 -> 12 -> 32 input block on fp32 NCHW input.  Each input value travels as a bf16 pair hi + lo (K = 16 holds
  // 6 hi, 6 lo, two constant-one channels and 2 zeros; W1 repeats its 6 columns), so the first 1x1 conv sees the fp32
  // input to ~16 bits.  The hidden tensor is padded 12 -> 16 channels; channels 12 and 13 are the constant 1 inside the
  // image all the way down, and ALL THREE folded-BN biases ride on them as hi + lo bf16 pairs in the weight tiles (exact
  // to ~2^-17): the drains and the epilogue only round, apply LeakyReLU on the packed values and store.  A zero A1 row
  // (outside the image) gives an exactly zero hidden row, which is the zero padding the depthwise conv needs.
  static constexpr int CH = CH_, NG = CH / 16, WW = SW + 2, S = W / SW, HP = W + 2;
  static constexpr int DT = CH >= 64 ? 256 : 128;          // drain threads that take part (32 or 16 columns each)
  static constexpr int OV = (2 * WW + 2 + 7) & ~7;          // rows a shifted view reads beyond its tile
  static constexpr int oA1 = 0;
  static constexpr int oHID = oA1 + 2 * kTileB;
  static constexpr int oA2 = oHID + 3 * kTileB + OV * 128;
  static constexpr int oW1 = oA2 + 2 * kTileB;
  static constexpr int oW2 = oW1 + CH * 128;
  static constexpr int oWD = oW2 + COUT * 128;              // 9 taps x [16 rows x 128 B]: diagonal blocks, group g at +32g
  static constexpr int oMETA = oWD + 9 * 2048;
  // INC: fp32 input planes staged by cp.async, kStageD tiles per producer group: [group][slot][6 planes + flags][128]
  static constexpr int kStageD = 4, kStageSlotB = 7 * 512;
  static constexpr int oSTG = oMETA + 4 * 128;
  static constexpr int oBAR = oSTG + ((INC && !STRIP_TC_INC_LDG) ? 2 * kStageD * kStageSlotB : 0);
  static constexpr int kSmem = oBAR + 512 + 1024;
  static constexpr uint32_t kWeightBytes = (INC ? 0 : CH * 128) + (INC ? 4096 : COUT * 128);
  static_assert(CH == 64 || (INC && CH == 16 && CIN == 16), "one 64-channel chunk of hidden channels (or the input block)");
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

// Position q = jrow * WW + hx of the CTA's strip raster -> (jrow, hx) and (frame-strip bs, padded row hy) of global row
// G0 + jrow; next() advances q by STEP without divisions.
template <int WW, int HP, int STEP>
struct PosIter {
  int jrow, hx, hy, bs;
  __device__ __forceinline__ void init(int q, int G0) {
    jrow = q / WW;
    hx = q - jrow * WW;
    const int G = G0 + jrow;
    bs = G / HP;
    hy = G - bs * HP;
  }
  __device__ __forceinline__ void next() {
    constexpr int DJ = STEP / WW, DX = STEP % WW;
    static_assert(DJ + 1 < HP, "at most one wrap per step");
    hx += DX;
    int dj = DJ;
    if (hx >= WW) {
      hx -= WW;
      ++dj;
    }
    jrow += dj;
    hy += dj;
    if (hy >= HP) {
      hy -= HP;
      ++bs;
    }
  }
};

// 8 fp32 accumulator columns -> 4 packed bf16x2 with LeakyReLU on the packed values (biases already inside the accumulator)
__device__ __forceinline__ void leaky8_packed(const uint32_t* acc, const __nv_bfloat162 kslope, uint32_t* o) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
    v = __hmax2(v, __hmul2(v, kslope));
    o[j] = *reinterpret_cast<uint32_t*>(&v);
  }
}

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
__global__ void __launch_bounds__(C::kThreads, 1) strip_tc_kernel(const __grid_constant__ StripArgs p) {
  constexpr int CIN = C::CIN, COUT = C::COUT, W = C::W, SW = C::SW, CH = C::CH, WW = C::WW, S = C::S, HP = C::HP,
                H = C::W, OV = C::OV, NG = C::NG, DT = C::DT;
  constexpr bool RES = C::RES, INC = C::INC, EG2 = C::EG2;
  constexpr int kThreads = C::kThreads;
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
      mbar_init(bar(B_D1FREE + i), DT);
      mbar_init(bar(B_DWFULL + i), 1);
      mbar_init(bar(B_DWFREE + i), DT);
      mbar_init(bar(B_A2FULL + i), DT);
      mbar_init(bar(B_A2FREE + i), 1);
      mbar_init(bar(B_D2FULL + i), 1);
      mbar_init(bar(B_D2FREE + i), 128);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(bar(B_HIDFULL + i), DT);
      mbar_init(bar(B_HIDFREE + i), 1);
    }
    fence_mbar_init();
    mbar_arrive_expect_tx(bar(B_W), C::kWeightBytes);
    if constexpr (!INC) bulk_g2s(sW1, p.W1, CH * 128, bar(B_W));
    bulk_g2s(sW2, p.W2, INC ? 4096 : COUT * 128, bar(B_W));   // INC: the packed pw2 tile (32 rows, K padded to 64)
  }
  if (warp == kIssWarp0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // diagonal tap blocks: tap t9 -> [16 rows n][64 k] bf16, SWIZZLE_128B image; element (n, 16g + n) = w[t9][16g + n]
  for (int i = tid; i < 9 * 2048 / 16; i += kThreads) sm.at<uint4>(sWD + i * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if constexpr (INC) {
    for (int i = tid; i < 16 * 128 / 16; i += kThreads) sm.at<uint4>(sW1 + i * 16) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if constexpr (!INC) {
    if (tid < 9 * 64) {
      const int t9 = tid >> 6, ch = tid & 63, g = ch >> 4, n = ch & 15;
      // wdp: bf16 [CH/8][10][8]
      const __nv_bfloat16 wv = reinterpret_cast<const __nv_bfloat16*>(p.wdp)[((ch >> 3) * 10 + t9) * 8 + (ch & 7)];
      const int k = 16 * g + n;   // column of row n
      sm.at<__nv_bfloat16>(sWD + t9 * 2048 + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = wv;
    }
  } else {
    auto put = [&](uint32_t tile, int n, int k, float v) {   // element (row n, column k) of a SWIZZLE_128B K-major tile
      sm.at<__nv_bfloat16>(tile + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
    };
    auto lo_part = [](float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); };
    if (tid < 9 * 12) {   // diagonal tap blocks from the fp32 taps
      const int t9 = tid / 12, n = tid - t9 * 12;
      put(sWD + t9 * 2048, n, n, p.inc_wd[t9 * 12 + n]);
    } else if (tid >= 128 && tid < 128 + 12 * 12) {   // W1 tile: row n = hidden channel, columns 0..5 (hi) and 6..11 (lo)
      const int i = tid - 128, n = i / 12, k = i - n * 12;
      put(sW1, n, k, p.inc_w1[n * 6 + (k < 6 ? k : k - 6)]);
    } else if (tid >= 288 && tid < 288 + 12) {   // biases b1 / bd on the constant-one channels 12 (hi) and 13 (lo)
      const int n = tid - 288;
      put(sW1, n, 12, p.b1[n]);
      put(sW1, n, 13, lo_part(p.b1[n]));
      put(sWD + 4 * 2048, n, 12, p.inc_bd[n]);
      put(sWD + 4 * 2048, n, 13, lo_part(p.inc_bd[n]));
    } else if (tid == 300 || tid == 301) {   // the ones propagate: hidden[12 + i] = 1 . 1, DW[12 + i] = hidden[12 + i] . 1
      const int n = 12 + (tid - 300);
      put(sW1, n, 12, 1.f);
      put(sWD + 4 * 2048, n, n, 1.f);
    } else if (tid >= 320 && tid < 320 + 32) {   // b2 into columns 12 / 13 of the pw2 tile once the bulk copy has landed
      const int n = tid - 320;
      mbar_wait(bar(B_W), 0);
      put(sW2, n, 12, p.b2[n]);
      put(sW2, n, 13, lo_part(p.b2[n]));
    }
  }
  fence_proxy_async();   // generic-proxy writes of WD -> async proxy (tcgen05.mma reads)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tD1 = tmem, tDW = tmem + 128, tD2 = tmem + 256;

  const long long NGR = (long long)p.batch * S * HP;   // rows of the global padded-row list
  const int G0 = (int)(NGR * blockIdx.x / gridDim.x), G1 = (int)(NGR * (blockIdx.x + 1) / gridDim.x);
  const int ntop = G1 - G0, nrows = ntop + 2;
  const int NT = (nrows * WW + 127) >> 7;
  const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);

  long long tmark = STRIP_DBG && p.dbg ? clock64() : 0;
  const bool timed = STRIP_DBG && p.dbg && lane == 0 &&
                     (warp == kProdWarp0 || warp == kDrain1Warp0 || warp == kDrain2Warp0 || warp == kEpiWarp0 ||
                      (warp >= kIssWarp0 && warp < kIssWarp0 + 3));
  unsigned long long tacc[23] = {};   // constant indices only: lives in registers, flushed once at the end of each role
  auto T = [&](int slot) {
    if (STRIP_DBG && timed) {
      const long long now = clock64();
      tacc[slot] += (unsigned long long)(now - tmark);
      tmark = now;
    }
  };
#define STRIP_TC_FLUSH(LO, HI)                                        \
  if (STRIP_DBG && timed) {                                           \
    _Pragma("unroll") for (int i_ = LO; i_ < HI; ++i_) atomicAdd(p.dbg + i_, tacc[i_]); \
  }

  if (warp >= kIssWarp0 && warp < kIssWarp0 + 4) {
    if (warp == kIssWarp0) {
      // ----- first 1x1 conv
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, CH);
      mbar_wait_sleep<kSleepNs>(bar(B_W), 0);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1;
        mbar_wait_sleep<kSleepNs>(bar(B_A1FULL + s), ph);
        T(4);
        mbar_wait_sleep<kSleepNs>(bar(B_D1FREE + s), ph ^ 1);
        T(5);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(sA1 + s * kTileB), bd = umma_desc_sw128(sW1);
#pragma unroll
          for (int ks = 0; ks < CIN / 16; ++ks) umma_bf16(tD1 + s * 64, ad + 2 * ks, bd + 2 * ks, idesc1, ks != 0);
          umma_commit(bar(B_D1FULL + s));
          if (!STRIP_TC_ONE_COMMIT) umma_commit(bar(B_A1FREE + s));
        }
        __syncwarp();
        T(6);
      }
      STRIP_TC_FLUSH(4, 7)
    } else if (warp == kIssWarp0 + 1) {
      // ----- depthwise 3x3 as 36 shifted N16 MMAs per tile
      constexpr uint32_t idesc16 = umma_idesc_bf16(128, 16);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1, hs = t % 3;
        mbar_wait_sleep<kSleepNs>(bar(B_HIDFULL + hs), (t / 3) & 1);
        T(10);
        if (t + 1 < NT) mbar_wait_sleep<kSleepNs>(bar(B_HIDFULL + (t + 1) % 3), ((t + 1) / 3) & 1);
        T(11);
        mbar_wait_sleep<kSleepNs>(bar(B_DWFREE + s), ph ^ 1);
        T(12);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a0 = umma_desc_sw128(sHID + hs * kTileB);
          const uint64_t b0 = umma_desc_sw128(sWD);
#pragma unroll
          for (int t9 = 0; t9 < ((STRIP_EXP & 1) ? 1 : 9); ++t9) {
            const uint64_t ad = a0 + (uint64_t)(((t9 / 3) * WW + (t9 % 3)) * 128 >> 4);   // start address += shift rows
            const uint64_t bd = b0 + (uint64_t)(t9 * 2048 >> 4);
#pragma unroll
            for (int g = 0; g < NG; ++g) umma_bf16(tDW + s * 64 + g * 16, ad + 2 * g, bd + 2 * g, idesc16, t9 != 0);
          }
          umma_commit(bar(B_DWFULL + s));
          if (!STRIP_TC_ONE_COMMIT) umma_commit(bar(B_HIDFREE + hs));
        }
        __syncwarp();
        T(13);
      }
      STRIP_TC_FLUSH(10, 14)
    } else if (warp == kIssWarp0 + 2) {
      // ----- second 1x1 conv
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, COUT);
      mbar_wait_sleep<kSleepNs>(bar(B_W), 0);
      for (int t = 0; t < NT; ++t) {
        const int s = t & 1, ph = (t >> 1) & 1;
        mbar_wait_sleep<kSleepNs>(bar(B_A2FULL + s), ph);
        T(17);
        mbar_wait_sleep<kSleepNs>(bar(B_D2FREE + s), ph ^ 1);
        T(18);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(sA2 + s * kTileB), bd = umma_desc_sw128(sW2);
#pragma unroll
          for (int ks = 0; ks < CH / 16; ++ks) umma_bf16(tD2 + s * COUT, ad + 2 * ks, bd + 2 * ks, idesc2, ks != 0);
          if (!STRIP_TC_ONE_COMMIT) umma_commit(bar(B_A2FREE + s));
          umma_commit(bar(B_D2FULL + s));
        }
        __syncwarp();
        T(19);
      }
      STRIP_TC_FLUSH(17, 20)
    }
  } else if ((warp >= kProdWarp0 && warp < kProdWarp0 + 4) ||
             (INC && warp >= kDrain1Warp0 + 4 && warp < kDrain1Warp0 + 8)) {
    // =========================================== producers: A1[t] <- global ====================================
    pdl_wait();
    const int r = (warp & 3) * 32 + lane;
    const uint32_t r7 = r & 7;
    if constexpr (INC) {
      // Two producer groups (the second one is the idle half of drain 1), tiles alternate between them.  The fp32 NCHW
      // planes are staged through shared memory with 4-byte cp.async, kStageD - 1 own tiles (6 KB per group) in flight;
      // every thread reads back only what it copied itself, so the staging needs no barrier.
      const int pg = warp >= kDrain1Warp0 ? 1 : 0;
      const int n_own = (NT - pg + 1) >> 1;
      PosIter<WW, HP, 256> pos;
      pos.init(pg * 128 + r, G0);
#if STRIP_TC_INC_LDG
      // registers one own tile (two tiles of the CTA, ~2 us) ahead: the loads of tile i+1 are issued before tile i is
      // converted and stored, and are first touched an iteration later
      float nv[6];
      uint32_t nones;
      auto fetch = [&]() {
        const int b = pos.bs / S, st = pos.bs - b * S;
        const int y = pos.hy - 1, x = st * SW + pos.hx - 1;
        const bool in = pos.jrow < nrows && pos.bs < p.batch * S && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        const float* src = p.x_nchw + (in ? (size_t)b * 6 * (H * W) + y * W + x : 0);
#pragma unroll
        for (int c = 0; c < 6; ++c) nv[c] = in ? __ldg(src + c * (H * W)) : 0.f;
        nones = in ? 0x3F803F80u : 0u;   // channels 12, 13: bf16 1.0 inside the image
        pos.next();
      };
      fetch();
      for (int i = 0; i < n_own; ++i) {
        float v[6], lo[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = nv[c];
        const uint32_t ones = nones;
        T(1);
        if (i + 1 < n_own) fetch();
        T(0);
        const int t = pg + 2 * i, s = t & 1;
#pragma unroll
        for (int c = 0; c < 6; ++c) lo[c] = v[c] - __bfloat162float(__float2bfloat16_rn(v[c]));
#else
      constexpr int D = C::kStageD;
      const uint32_t stg = base + C::oSTG + pg * D * C::kStageSlotB + r * 4;
      for (int i = 0; i < n_own + D - 1; ++i) {
        if (i < n_own) {
          const int b = pos.bs / S, st = pos.bs - b * S;
          const int y = pos.hy - 1, x = st * SW + pos.hx - 1;
          const bool in = pos.jrow < nrows && pos.bs < p.batch * S && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
          const float* src = p.x_nchw + (in ? (size_t)b * 6 * (H * W) + y * W + x : 0);
          const uint32_t d = stg + (i % D) * C::kStageSlotB;
#pragma unroll
          for (int c = 0; c < 6; ++c) cp_async4(d + c * 512, src + c * (H * W), in);
          sm.at<uint32_t>(d + 6 * 512) = in ? 0x3F803F80u : 0u;   // channels 12, 13: bf16 1.0 inside the image
          pos.next();
        }
        cp_async_commit();
        T(0);
        const int u = i - (D - 1);
        if (u < 0) continue;
        cp_async_wait<D - 1>();
        T(1);
        const int t = pg + 2 * u, s = t & 1;
        const uint32_t d = stg + (u % D) * C::kStageSlotB;
        float v[6], lo[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          v[c] = sm.at<float>(d + c * 512);
          lo[c] = v[c] - __bfloat162float(__float2bfloat16_rn(v[c]));
        }
        const uint32_t ones = sm.at<uint32_t>(d + 6 * 512);
#endif
        T(3);
        if (t >= 2) mbar_wait_sleep<kSleepNs>(bar((STRIP_TC_ONE_COMMIT ? B_D1FULL : B_A1FREE) + s), ((t >> 1) & 1) ^ 1);
        T(2);
        const uint32_t a1 = sA1 + s * kTileB + r * 128;
        sm.at<uint4>(a1 + ((0 ^ r7) << 4)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(lo[0], lo[1]));
        sm.at<uint4>(a1 + ((1 ^ r7) << 4)) = make_uint4(pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), ones, 0u);
        fence_proxy_async();
        mbar_arrive(bar(B_A1FULL + s));
        T(3);
      }
      STRIP_TC_FLUSH(0, 4)
    } else {
    PosIter<WW, HP, 128> pos;
    pos.init(r, G0);
    for (int t = 0; t <= NT; ++t) {
      if (t < NT) {
        const int s = t & 1;
        if (t >= 2) mbar_wait_sleep<kSleepNs>(bar((STRIP_TC_ONE_COMMIT ? B_D1FULL : B_A1FREE) + s), ((t >> 1) & 1) ^ 1);
        T(2);
        const int b = pos.bs / S, st = pos.bs - b * S;
        const int y = pos.hy - 1, x = st * SW + pos.hx - 1;
        const bool inside =
            pos.jrow < nrows && pos.bs < p.batch * S && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        pos.next();
        sm.at<uint8_t>(sMETA + (t & 3) * 128 + r) = inside ? 1 : 0;
        const uint32_t a1 = sA1 + s * kTileB + r * 128;
        const size_t pix = inside ? ((size_t)b * H + y) * W + x : 0;
        const __nv_bfloat16* src = p.in + pix * CIN;
#pragma unroll
        for (int c = 0; c < CIN / 8; ++c) cp_async16(a1 + ((c ^ r7) << 4), src + c * 8, inside);
        cp_async_commit();
        T(0);
      }
      if (t >= 1) {   // tile t-1 has landed once at most the newest group is still in flight
        if (t < NT) cp_async_wait<1>(); else cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(bar(B_A1FULL + ((t - 1) & 1)));
        T(1);
      }
    }
    STRIP_TC_FLUSH(0, 4)
    }
  } else if (warp >= kDrain1Warp0 && warp < kDrain1Warp0 + (INC ? 4 : 8)) {
    // =========================================== drain 1: D1 -> HID ring =======================================
    const int dwp = warp - kDrain1Warp0, lg = dwp & 3, hw = dwp >> 2;   // TMEM lane quarter, column half (32 columns)
    const int row = lg * 32 + lane;
    const uint32_t r7 = row & 7;
    for (int t = 0; t < NT; ++t) {   // INC: 16 hidden columns, one warp per lane quarter
      const int b = t & 1, hs = t % 3;
      mbar_wait_sleep<kSleepNs>(bar(B_D1FULL + b), (t >> 1) & 1);
      T(7);
      if (t >= 3) {   // dw(t - 3) has completed: it was the last reader of this ring slot
        if (STRIP_TC_ONE_COMMIT) mbar_wait_sleep<kSleepNs>(bar(B_DWFULL + ((t - 3) & 1)), ((t - 3) >> 1) & 1);
        else mbar_wait_sleep<kSleepNs>(bar(B_HIDFREE + hs), ((t / 3) & 1) ^ 1);
      }
      T(8);
      tc_fence_after();
      if constexpr (INC) {
        uint32_t acc[16];
        tmem_ld16(tD1 + b * 64 + ((uint32_t)(lg * 32) << 16), acc);
        tmem_ld_wait16(acc);
        tc_fence_before();
        mbar_arrive(bar(B_D1FREE + b));
        const uint32_t hid = sHID + hs * kTileB + row * 128;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint32_t o[4];
          leaky8_packed(acc + g8 * 8, kslope, o);   // bias and the zero rows outside the image come out of the MMA
          const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
          sm.at<uint4>(hid + (((uint32_t)g8 ^ r7) << 4)) = v;
          if (hs == 0 && row < OV) sm.at<uint4>(hid + 3 * kTileB + (((uint32_t)g8 ^ r7) << 4)) = v;
        }
        fence_proxy_async();
        mbar_arrive(bar(B_HIDFULL + hs));
        T(9);
        continue;
      }
      const bool inside = sm.at<uint8_t>(sMETA + (t & 3) * 128 + row) != 0;
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
      T(9);
    }
    STRIP_TC_FLUSH(7, 10)
  } else if (warp >= kDrain2Warp0 && warp < kDrain2Warp0 + (INC ? 4 : 8)) {
    // =========================================== drain 2: DW -> A2 ==============================================
    const int dwp = warp - kDrain2Warp0, lg = dwp & 3, hw = dwp >> 2;
    const int row = lg * 32 + lane;
    const uint32_t r7 = row & 7;
    // folded-BN bias of the depthwise conv for this thread's 32 channels: packed bf16 pairs as stored in wdp (row 9 of
    // every 8-channel entry); added after the fp32 accumulator is rounded to bf16 (the taps themselves are bf16)
    __nv_bfloat162 bd2[16];
#pragma unroll
    for (int i = 0; i < (INC ? 0 : 4); ++i) {
      const uint4 b8 = __ldg(reinterpret_cast<const uint4*>(p.wdp + (size_t)(hw * 4 + i) * 160 + 9 * 16));
      const uint32_t* pb = &b8.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) bd2[i * 4 + j] = *reinterpret_cast<const __nv_bfloat162*>(&pb[j]);
    }
    for (int t = 0; t < NT; ++t) {
      const int b = t & 1;
      mbar_wait_sleep<kSleepNs>(bar(B_DWFULL + b), (t >> 1) & 1);
      T(14);
      if (t >= 2) mbar_wait_sleep<kSleepNs>(bar((STRIP_TC_ONE_COMMIT ? B_D2FULL : B_A2FREE) + b), ((t >> 1) & 1) ^ 1);
      T(15);
      tc_fence_after();
      if constexpr (INC) {
        uint32_t acc[16];
        tmem_ld16(tDW + b * 64 + ((uint32_t)(lg * 32) << 16), acc);
        tmem_ld_wait16(acc);
        tc_fence_before();
        mbar_arrive(bar(B_DWFREE + b));
        const uint32_t a2 = sA2 + b * kTileB + row * 128;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint32_t o[4];
          leaky8_packed(acc + g8 * 8, kslope, o);
          sm.at<uint4>(a2 + (((uint32_t)g8 ^ r7) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        mbar_arrive(bar(B_A2FULL + b));
        T(16);
        continue;
      }
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
      T(16);
    }
    STRIP_TC_FLUSH(14, 17)
  } else if ((warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) || (INC && warp >= kDrain2Warp0 + 4 && warp < kDrain2Warp0 + 8) ||
             (EG2 && warp >= 28)) {
    // =========================================== epilogue: D2 -> global ========================================
    // INC: two epilogue groups (the second one is the idle half of drain 2), one per D2 slot
    pdl_wait();
    const int row = (warp & 3) * 32 + lane;
    const bool fuse_outc = !INC && COUT == 32 && p.final_out != nullptr;   // last decoder block: the output head runs here
    const int t0 = ((INC && warp >= kDrain2Warp0) || (EG2 && warp >= 28)) ? 1 : 0;
    PosIter<WW, HP, (INC || EG2) ? 256 : 128> pos;
    pos.init(t0 * 128 + row, G0);
    for (int t = t0; t < NT; t += ((INC || EG2) ? 2 : 1)) {
      const int s = t & 1;
      const int jrow = pos.jrow, hx = pos.hx, bs = pos.bs, hy = pos.hy;
      pos.next();
      const int b = bs / S, st = bs - b * S;
      const bool valid = jrow < ntop && hy < H && hx < SW && bs < p.batch * S;
      const size_t pix = valid ? ((size_t)b * H + hy) * W + st * SW + hx : 0;
      if constexpr (INC) {
        T(20);
        mbar_wait_sleep<kSleepNs>(bar(B_D2FULL + s), (t >> 1) & 1);
        T(21);
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32(tD2 + s * COUT + ((uint32_t)((warp & 3) * 32) << 16), acc);
        tmem_ld_wait32(acc);
        tc_fence_before();
        mbar_arrive(bar(B_D2FREE + s));
        if (valid && !((STRIP_EXP & 8) && acc[0] != 0x12345u)) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            uint32_t o[4];
            leaky8_packed(acc + g8 * 8, kslope, o);
            *reinterpret_cast<uint4*>(p.out + pix * COUT + g8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        T(22);
        continue;
      }
      uint4 rr[RES ? COUT / 8 : 1];
      if constexpr (RES) {
#pragma unroll
        for (int i = 0; i < COUT / 8; ++i)
          rr[i] = (valid && !(STRIP_EXP & 8)) ? __ldg(reinterpret_cast<const uint4*>(p.in + pix * CIN + i * 8))
                                              : make_uint4(0, 0, 0, 0);
      }
      T(20);
      mbar_wait_sleep<kSleepNs>(bar(B_D2FULL + s), (t >> 1) & 1);
      T(21);
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
      T(22);
    }
    STRIP_TC_FLUSH(20, 23)
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
  const long long rows = (long long)a.batch * C::S * C::HP;
  long long grid = rows / 6;
  if (grid < 1) grid = 1;
  if (grid > num_sms) grid = num_sms;
  return (int)launch_pdl(kfn, dim3((unsigned)grid), dim3(C::kThreads), C::kSmem, st, a);
}

}  // namespace

#ifndef STRIP_TC_EG2
#define STRIP_TC_EG2 false   // measured: up4.1 127.8 against 125.3 us with a second epilogue group in warps 28..31
#endif
#define STRIP_TC_CASES(X)                                                \
  X(32, 32, 160, 40, true, STRIP_TC_EG2)    /* up4.1 */                  \
  X(32, 32, 80, 40, true, STRIP_TC_EG2)     /* up3.1 */                  \
  X(32, 64, 32, 32, false, false)           /* audio conv1 */

bool strip_tc_supported(int cin, int cout, int W, int stride, bool upcat, bool res) {
  if (stride != 1 || upcat) return false;
#define X(CIN_, COUT_, W_, SW_, R_, E_) \
  if (cin == CIN_ && cout == COUT_ && W == W_ && res == R_) return true;
  STRIP_TC_CASES(X)
#undef X
  return false;
}

int launch_strip_tc(const StripArgs& a, int cin, int cout, int W, int stride, bool upcat, bool res, int num_sms,
                    cudaStream_t st) {
  if (stride != 1 || upcat) return -1;
#define X(CIN_, COUT_, W_, SW_, R_, E_)                                   \
  if (cin == CIN_ && cout == COUT_ && W == W_ && res == R_)               \
    return launch_t<TCfg<CIN_, COUT_, W_, SW_, R_, 2 * CIN_, false, E_>>(a, num_sms, st);
  STRIP_TC_CASES(X)
#undef X
  return -1;
}

int launch_strip_inc(const StripArgs& a, int num_sms, cudaStream_t st) {
  return launch_t<TCfg<16, 32, 160, 40, false, 16, true>>(a, num_sms, st);
}

}  // namespace casync
