// Fused InvertedResidual (module/unet.py:8-40) for the high-resolution, HBM-bound blocks:
//   pw1 (tcgen05) -> BN+leaky -> depthwise 3x3 (CUDA cores, from shared memory) -> BN+leaky -> pw2 (tcgen05)
//   -> BN+leaky (+ skip)            -- the x2-expanded hidden tensor never leaves the SM.
//
// One persistent, warp-specialised CTA per SM walks over 2-D patches of one frame.  A patch is a window of
// 16 x (8*TILES) HIDDEN pixels (TILES 128-row UMMA tiles): stride 1 -> 14 x (8*TILES-2) outputs, stride 2 ->
// 7 x ((8*TILES-1)/2).  The hidden channels are processed in chunks of 64 ("units" = patch x chunk) through a
// double-buffered pipeline linked by mbarriers:
//
//   producers (4 warps) A1[patch] <- global (cp.async rows; decoder: bilinear x2 of the low-res tensor computed
//                                    on the fly for the first Cin/2 channels + skip rows, module/unet.py:90-96)
//   issuer  (1 thread) D1[u]      =  A1 . W1[chunk]^T                     tcgen05.mma  -> TMEM (fp32)
//   group B            HID[u]     =  leaky(D1 + b1) as bf16, UMMA-swizzled smem; 0 outside the image (the
//                                    depthwise conv zero-pads the hidden tensor, module/unet.py:21-27)
//   group A (8 warps)  A2[u]      =  leaky(dw3x3(HID) + bd)               packed bf16x2 FMAs on CUDA cores
//   issuer             D2[patch] +=  A2 . W2[:, chunk]^T                  tcgen05.mma  -> TMEM (fp32)
//   group B            out        =  leaky(D2 + b2) (+ x)  -> global NHWC bf16
//
// so TMEM drains, global-memory latency, CUDA-core depthwise math and both MMA groups of neighbouring units
// overlap.  Weights (W1/W2 in their packed swizzled image, depthwise taps, biases) are fetched once per CTA with
// bulk-async copies and stay resident in shared memory.
#include "fused_ir.cuh"

#include <cuda_bf16.h>

#ifndef FUSED_OPT_DRAIN2
#define FUSED_OPT_DRAIN2 0
#endif
#ifndef FUSED_OPT_RESPF
#define FUSED_OPT_RESPF 1
#endif
#ifndef FUSED_OPT_DWPF
#define FUSED_OPT_DWPF 1
#endif

namespace casync {

namespace {

constexpr int kGroup = 256;                 // threads per compute group (8 warps)
constexpr int kProd = 96;                   // A1 producer threads (3 warps): 20 warps per CTA -> 96 registers per thread
constexpr int kIssuerWarp = (2 * kGroup + kProd) / 32;
constexpr int kThreads = 2 * kGroup + kProd + 32;   // group B + group A + producers + issuer warp
constexpr int kTile = 128 * 128;            // bytes of one 128-row x 128 B swizzled tile

// WS ("weight streaming"): W1 / W2 do not stay resident -- the 64-row slice of W1 and the k-block of W2 that one unit
// (patch x hidden chunk) needs are fetched into two-deep rings by the MMA issuer, one unit ahead, so blocks with
// CIN >= 128 (down2.1: 128 KB of weights, up2.0: 320 KB) fit.  The depthwise taps then come pre-packed as bf16
// ([CH/8][10][8], the layout of the stand-alone depthwise kernel) to save 13 bytes of shared memory per channel.
template <int CIN, int COUT, int STRIDE, int TILES, int A1BUFS, bool WS = false>
struct FCfg {
  static constexpr int CH = 2 * CIN, NC = CH / 64, KB1 = (CIN + 63) / 64;
  static constexpr int WIN_H = 8 * TILES;
  static constexpr int TOH = STRIDE == 1 ? WIN_H - 2 : (WIN_H - 1) / 2;
  static constexpr int TOW = STRIDE == 1 ? 14 : 7;
  static constexpr int NOUT = TOH * TOW;
  static constexpr int A2T = (NOUT + 127) / 128;
  static constexpr int kA1Buf = KB1 * TILES * kTile, kHidBuf = TILES * kTile, kA2Buf = A2T * kTile;
  static constexpr int oA1 = 0;
  static constexpr int oHID = oA1 + A1BUFS * kA1Buf;
  static constexpr int oA2 = oHID + 2 * kHidBuf;
  static constexpr int kW1Chunk = KB1 * 64 * 128, kW2Chunk = COUT * 128;   // what one unit reads of W1 / W2
  static constexpr int oW1 = oA2 + 2 * kA2Buf;
  static constexpr int oW2 = oW1 + (WS ? 2 * kW1Chunk : KB1 * CH * 128);
  static constexpr int oWD = oW2 + (WS ? 2 * kW2Chunk : NC * COUT * 128);
  static constexpr int oB1 = oWD + (WS ? CH * 20 : 9 * CH * 4);
  static constexpr int oBD = oB1 + CH * 4;
  static constexpr int oB2 = oBD + (WS ? 0 : CH * 4);
  static constexpr int oBAR = oB2 + COUT * 4;
  static constexpr int kSmem = oBAR + 256 + 1024;
  static constexpr uint32_t kWeightBytes =
      WS ? CH * 20 + CH * 4 + COUT * 4 : KB1 * CH * 128 + NC * COUT * 128 + 9 * CH * 4 + 2 * CH * 4 + COUT * 4;
  static constexpr int kTmemD2 = 2 * TILES * 64;   // D1 double buffer first, then D2
  static_assert(kTmemD2 + A2T * COUT <= 512, "TMEM overflow");
  static_assert(kSmem <= 232448, "shared memory overflow");
};

enum Bar : int {
  B_W = 0, B_A1FULL = 1, B_A1FREE = 3, B_D1FULL = 5, B_D1FREE = 7, B_HIDFULL = 9, B_HIDFREE = 11, B_A2FULL = 13,
  B_A2FREE = 15, B_D2FULL = 17, B_D2FREE = 18, B_W1FULL = 19, B_W1FREE = 21, B_W2FULL = 23, B_W2FREE = 25, B_COUNT = 27
};

// Shared-memory accesses of the CUDA-core phases are ordinary C++ loads/stores (through the generic pointer of
// the dynamic smem array, which the compiler resolves to LDS/STS) so that ptxas can batch and reorder them
// freely between the mbarrier operations; an `asm volatile(... "memory")` per access serialised every phase.
struct SmemView {
  uint8_t* g;      // generic pointer of the 1024-aligned base
  uint32_t base;   // its shared-window address
  template <class T>
  __device__ __forceinline__ T& at(uint32_t addr) const { return *reinterpret_cast<T*>(g + (addr - base)); }
};

struct Patch {
  int b, OY0, OX0, GY0, GX0;
};

template <int CIN, int COUT, int STRIDE, bool UPCAT, bool RES, int TILES, int A1BUFS, bool WS>
__global__ void __launch_bounds__(kThreads, 1) fused_ir_kernel(const FusedArgs p) {
  using C = FCfg<CIN, COUT, STRIDE, TILES, A1BUFS, WS>;
  constexpr int CH = C::CH, NC = C::NC, KB1 = C::KB1, TOH = C::TOH, TOW = C::TOW, NOUT = C::NOUT, A2T = C::A2T;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const SmemView sm{smem_raw + (base - smem_u32(smem_raw)), base};
  auto lds64 = [&](uint32_t addr) { return sm.at<uint2>(addr); };
  auto lds_f4 = [&](uint32_t addr) { return sm.at<float4>(addr); };
  auto sts64 = [&](uint32_t addr, uint32_t a, uint32_t b) { sm.at<uint2>(addr) = make_uint2(a, b); };
  auto sts128 = [&](uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    sm.at<uint4>(addr) = make_uint4(a, b, c, d);
  };
  const uint32_t sA1 = base + C::oA1, sHID = base + C::oHID, sA2 = base + C::oA2, sW1 = base + C::oW1,
                 sW2 = base + C::oW2, sWD = base + C::oWD, sB1 = base + C::oB1, sBD = base + C::oBD,
                 sB2 = base + C::oB2, sBAR = base + C::oBAR;
  auto bar = [&](int i) { return sBAR + 8u * i; };
  const uint32_t tmem_slot = sBAR + 8u * B_COUNT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();

  if (tid == 0) {
    mbar_init(bar(B_W), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_A1FULL + i), kProd);
      mbar_init(bar(B_A1FREE + i), 1);
      mbar_init(bar(B_D1FULL + i), 1);
      mbar_init(bar(B_D1FREE + i), kGroup);
      mbar_init(bar(B_HIDFULL + i), kGroup);
      mbar_init(bar(B_HIDFREE + i), kGroup);
      mbar_init(bar(B_A2FULL + i), kGroup);
      mbar_init(bar(B_A2FREE + i), 1);
    }
    mbar_init(bar(B_D2FULL), 1);
    mbar_init(bar(B_D2FREE), kGroup);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_W1FULL + i), 1);
      mbar_init(bar(B_W1FREE + i), 1);
      mbar_init(bar(B_W2FULL + i), 1);
      mbar_init(bar(B_W2FREE + i), 1);
    }
    fence_mbar_init();
    mbar_arrive_expect_tx(bar(B_W), C::kWeightBytes);
    if constexpr (WS) {
      bulk_g2s(sWD, p.wdp, CH * 20, bar(B_W));
    } else {
      bulk_g2s(sW1, p.W1, KB1 * CH * 128, bar(B_W));
      bulk_g2s(sW2, p.W2, NC * COUT * 128, bar(B_W));
      bulk_g2s(sWD, p.wd, 9 * CH * 4, bar(B_W));
      bulk_g2s(sBD, p.bd, CH * 4, bar(B_W));
    }
    bulk_g2s(sB1, p.b1, CH * 4, bar(B_W));
    bulk_g2s(sB2, p.b2, COUT * 4, bar(B_W));
  }
  if (warp == kIssuerWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tD2 = tmem + C::kTmemD2;
  mbar_wait(bar(B_W), 0);
  pdl_wait();   // weights (constant) are already on their way; activations only after the previous kernel is done

  const int W = p.W, Wo = W / STRIDE;
  const int PDX = (Wo + TOW - 1) / TOW, PDY = (Wo + TOH - 1) / TOH;
  const int NP = p.batch * PDX * PDY;
  const int first = blockIdx.x, step = gridDim.x;
  const int n_mine = first < NP ? (NP - first + step - 1) / step : 0;
  auto decode = [&](int pi) {
    const int patch = first + pi * step;
    Patch q;
    q.b = patch / (PDX * PDY);
    const int pr = patch - q.b * PDX * PDY, py = pr / PDX, px = pr - py * PDX;
    q.OY0 = py * TOH;
    q.OX0 = px * TOW;
    q.GY0 = q.OY0 * STRIDE - 1;   // global coords of hidden-window pixel (0,0)
    q.GX0 = q.OX0 * STRIDE - 1;
    return q;
  };
  const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);
  // developer timing (CASYNC_PHASE_DBG): one thread per role accumulates cycles per activity slot
  long long tmark = p.dbg ? clock64() : 0;
  const bool timed = p.dbg && (tid == 0 || tid == kGroup || tid == 2 * kGroup || tid == kIssuerWarp * 32);
  auto T = [&](int slot) {
    if (timed) {
      const long long now = clock64();
      atomicAdd(p.dbg + slot, (unsigned long long)(now - tmark));
      tmark = now;
    }
  };

  if (warp == kIssuerWarp) {
    // =========================================== MMA issuer ====================================================
    // the whole (converged) warp walks the schedule and waits; one elected lane issues the tcgen05 instructions
    if (n_mine > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64), idesc2 = umma_idesc_bf16(128, COUT);
      // weight streaming: unit v's slice of W1 (needed by its first GEMM) and k-block of W2 (second GEMM) go to ring
      // slot v & 1; a slot is refilled once the MMAs of unit v - 2 that read it have completed (tcgen05.commit)
      auto load_w1 = [&](int v) {
        const int cv = v % NC, b = v & 1, k = v >> 1;
        mbar_wait(bar(B_W1FREE + b), (k & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(B_W1FULL + b), C::kW1Chunk);
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb)
            bulk_g2s(sW1 + b * C::kW1Chunk + kb * 64 * 128, p.W1 + ((size_t)kb * CH + cv * 64) * 128, 64 * 128, bar(B_W1FULL + b));
        }
        __syncwarp();
      };
      auto load_w2 = [&](int v) {
        const int cv = v % NC, b = v & 1, k = v >> 1;
        mbar_wait(bar(B_W2FREE + b), (k & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(B_W2FULL + b), C::kW2Chunk);
          bulk_g2s(sW2 + b * C::kW2Chunk, p.W2 + (size_t)cv * COUT * 128, C::kW2Chunk, bar(B_W2FULL + b));
        }
        __syncwarp();
      };
      auto gemm2 = [&](int v) {   // D2[patch] (+)= A2[v] . W2[:, chunk]^T
        const int pv = v / NC, cv = v - pv * NC, b = v & 1, k = v >> 1;
        mbar_wait(bar(B_A2FULL + b), k & 1);
        if (cv == 0) mbar_wait(bar(B_D2FREE), (pv & 1) ^ 1);
        if constexpr (WS) mbar_wait(bar(B_W2FULL + b), k & 1);
        T(3);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bd = umma_desc_sw128(WS ? sW2 + b * C::kW2Chunk : sW2 + cv * COUT * 128);
#pragma unroll
          for (int t = 0; t < A2T; ++t) {
            const uint64_t ad = umma_desc_sw128(sA2 + b * C::kA2Buf + t * kTile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(tD2 + t * COUT, ad + 2 * ks, bd + 2 * ks, idesc2, (cv | ks) != 0);
          }
          umma_commit(bar(B_A2FREE + b));
          if constexpr (WS) umma_commit(bar(B_W2FREE + b));
          if (cv == NC - 1) umma_commit(bar(B_D2FULL));
        }
        __syncwarp();
        T(4);
      };
      int u = 0;
      if constexpr (WS) load_w1(0);
      for (int pi = 0; pi < n_mine; ++pi) {
        const int ab = pi % A1BUFS, ak = pi / A1BUFS;
        mbar_wait(bar(B_A1FULL + ab), ak & 1);
        T(0);
        for (int c = 0; c < NC; ++c, ++u) {
          const int b = u & 1, k = u >> 1;
          if constexpr (WS) {
            if (u + 1 < n_mine * NC) load_w1(u + 1);   // one unit ahead (its slot was read by the first GEMM of unit u - 1)
            mbar_wait(bar(B_W1FULL + b), k & 1);
          }
          mbar_wait(bar(B_D1FREE + b), (k & 1) ^ 1);
          T(1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
#pragma unroll
              for (int kb = 0; kb < KB1; ++kb) {
                const uint64_t ad = umma_desc_sw128(sA1 + ab * C::kA1Buf + (kb * TILES + t) * kTile);
                const uint64_t bd = umma_desc_sw128(WS ? sW1 + b * C::kW1Chunk + kb * 64 * 128 : sW1 + (kb * CH + c * 64) * 128);
                constexpr int KS_ALL = (CIN + 15) / 16;
                const int ks_n = (KS_ALL - kb * 4) < 4 ? (KS_ALL - kb * 4) : 4;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < ks_n) umma_bf16(tmem + (b * TILES + t) * 64, ad + 2 * ks, bd + 2 * ks, idesc1, (kb | ks) != 0);
              }
            }
            umma_commit(bar(B_D1FULL + b));
            if constexpr (WS) umma_commit(bar(B_W1FREE + b));
            if (c == NC - 1) umma_commit(bar(B_A1FREE + ab));
          }
          __syncwarp();
          T(2);
          // W2 of this unit: needed by gemm2(u), one step from now; its slot was read by gemm2(u - 2), issued a whole
          // step ago (asking for it before issue1 made the issuer wait for that MMA group to retire)
          if constexpr (WS) load_w2(u);
          if (u > 0) gemm2(u - 1);
        }
      }
      gemm2(u - 1);
    }
  } else if (warp < 8) {
    // =========================================== group B: A1 producer, D1 drain, epilogue ======================
    const int lg = warp & 3, hw = warp >> 2;   // TMEM lane quarter, half (tile or column half)
    auto drain1 = [&](const Patch& q, int u, int c) {
      const int b = u & 1, k = u >> 1;
      T(15);
      mbar_wait(bar(B_D1FULL + b), k & 1);
      T(7);
      mbar_wait(bar(B_HIDFREE + b), (k & 1) ^ 1);
      T(8);
      tc_fence_after();
      constexpr int NCOL = TILES == 2 ? 64 : 32;   // columns per thread
      const int tile = TILES == 2 ? hw : 0, col0 = TILES == 2 ? 0 : hw * 32;
      const int row = tile * 128 + lg * 32 + lane;
      const int gy = q.GY0 + (row >> 4), gx = q.GX0 + (row & 15);
      const bool inside = gy >= 0 && gy < W && gx >= 0 && gx < W;
      const uint32_t hid = sHID + b * C::kHidBuf + tile * kTile + (row & 127) * 128;
      const uint32_t r7 = row & 7;
      const uint32_t tsrc = tmem + (b * TILES + tile) * 64 + col0 + ((uint32_t)(lg * 32) << 16);
      uint32_t accA[32], accB[32];
      tmem_ld32(tsrc, accA);
      tmem_ld_wait32(accA);
#if FUSED_OPT_DRAIN2
      if constexpr (NCOL == 64) tmem_ld32(tsrc + 32, accB);   // in flight while the first half is converted
#endif
      auto convert = [&](const uint32_t* acc, int cc) {
        // bias loads first, stores last: a shared-memory load cannot move above an earlier shared-memory store that
        // might alias, so loading the bias per 8-column group serialised the groups (LDS -> FADD -> ... -> STS -> LDS)
        float4 bq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bq[i] = lds_f4(sB1 + (c * 64 + col0 + cc + 4 * i) * 4);
        uint32_t o[4][4];
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const float bv[8] = {bq[2 * g8].x, bq[2 * g8].y, bq[2 * g8].z, bq[2 * g8].w,
                               bq[2 * g8 + 1].x, bq[2 * g8 + 1].y, bq[2 * g8 + 1].z, bq[2 * g8 + 1].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // bias in fp32, one rounding to bf16, LeakyReLU on the packed pair
            __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[g8 * 8 + 2 * j]) + bv[2 * j],
                                                     __uint_as_float(acc[g8 * 8 + 2 * j + 1]) + bv[2 * j + 1]);
            v = __hmax2(v, __hmul2(v, kslope));
            o[g8][j] = inside ? *reinterpret_cast<uint32_t*>(&v) : 0u;
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int col = col0 + cc + g8 * 8;
          sts128(hid + ((((uint32_t)col >> 3) ^ r7) << 4), o[g8][0], o[g8][1], o[g8][2], o[g8][3]);
        }
      };
      convert(accA, 0);
      if constexpr (NCOL == 64) {
#if !FUSED_OPT_DRAIN2
        tmem_ld32(tsrc + 32, accB);
#endif
        tmem_ld_wait32(accB);
        convert(accB, 32);
      }
      tc_fence_before();
      mbar_arrive(bar(B_D1FREE + b));
      mbar_arrive(bar(B_HIDFULL + b));
      T(9);
    };

    auto epilogue = [&](const Patch& q, int pi) {
      T(15);
      constexpr int NCOL = A2T == 2 ? COUT : COUT / 2;
      const int tile = A2T == 2 ? hw : 0, col0 = A2T == 2 ? 0 : hw * (COUT / 2);
      const int op = tile * 128 + lg * 32 + lane;
      const int oy = op / TOW, ox = op - oy * TOW;
      const int gy = q.OY0 + oy, gx = q.OX0 + ox;
      const bool valid = op < NOUT && gy < Wo && gx < Wo;
      const size_t opix = (size_t)(q.b * Wo + (valid ? gy : 0)) * Wo + (valid ? gx : 0);
      // the skip rows (block input at the output pixel) are fetched BEFORE waiting for the accumulator, so their
      // global-memory latency overlaps the depthwise / second GEMM of this patch
      uint4 rr[RES ? NCOL / 8 : 1];
#if FUSED_OPT_RESPF
      if constexpr (RES) {
#pragma unroll
        for (int i = 0; i < NCOL / 8; ++i)
          rr[i] = valid ? __ldg(reinterpret_cast<const uint4*>(p.in + opix * CIN + col0 + i * 8)) : make_uint4(0, 0, 0, 0);
      }
#endif
      mbar_wait(bar(B_D2FULL), pi & 1);
#if !FUSED_OPT_RESPF
      if constexpr (RES) {
#pragma unroll
        for (int i = 0; i < NCOL / 8; ++i)
          rr[i] = valid ? __ldg(reinterpret_cast<const uint4*>(p.in + opix * CIN + col0 + i * 8)) : make_uint4(0, 0, 0, 0);
      }
#endif
      T(10);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NCOL; cc += 16) {
        uint32_t acc[16];
        tmem_ld16(tD2 + tile * COUT + col0 + cc + ((uint32_t)(lg * 32) << 16), acc);
        tmem_ld_wait16(acc);
        if (valid) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const int col = col0 + cc + g8 * 8;
            const float4 ba = lds_f4(sB2 + col * 4), bb = lds_f4(sB2 + col * 4 + 16);
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] = __uint_as_float(acc[g8 * 8 + j]) + bv[j];
              v[j] = fmaxf(v[j], kLeaky * v[j]);
            }
            if constexpr (RES) {   // stride 1, CIN == COUT: skip = the block input at the same pixel
              const uint32_t* pr = &rr[(cc >> 3) + g8].x;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[2 * j] += bf16_lo(pr[j]);
                v[2 * j + 1] += bf16_hi(pr[j]);
              }
            }
            *reinterpret_cast<uint4*>(p.out + opix * p.ldo + col) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar(B_D2FREE));
      T(11);
    };

    if (n_mine > 0) {
      int u = 0;
      Patch prev{};
      for (int pi = 0; pi < n_mine; ++pi) {
        const Patch q = decode(pi);
        for (int c = 0; c < NC; ++c, ++u) {
          drain1(q, u, c);
          // the previous patch's epilogue is deferred until the next patch's first hidden chunk is in shared
          // memory, so the depthwise warps never wait for this group to finish an epilogue first
          if (c == 0 && pi > 0) epilogue(prev, pi - 1);
        }
        prev = q;
      }
      epilogue(prev, n_mine - 1);
    }
  } else if (warp >= 16) {
    // =========================================== producers: A1[patch] <- global ==================================
    auto produce_a1 = [&](int pi) {
      const Patch q = decode(pi);
      const int ab = pi % A1BUFS, ak = pi / A1BUFS;
      T(15);
      mbar_wait(bar(B_A1FREE + ab), (ak & 1) ^ 1);
      T(5);
      constexpr int TPR = 1;                        // one producer thread per window pixel ...
      constexpr int NCH = CIN / 8;                  // 16-byte chunks per pixel
      constexpr int sub = 0;
#pragma unroll 1
      for (int row = tid - 2 * kGroup; row < TILES * 128; row += kProd) {   // ... TILES pixels per thread
      const int gy = q.GY0 + (row >> 4), gx = q.GX0 + (row & 15);
      const bool inside = gy >= 0 && gy < W && gx >= 0 && gx < W;
      const uint32_t a1 = sA1 + ab * C::kA1Buf + (row >> 7) * kTile;
      const int r = row & 127;
      if constexpr (!UPCAT) {
        const __nv_bfloat16* src = p.in + ((size_t)(q.b * W + (inside ? gy : 0)) * W + (inside ? gx : 0)) * CIN;
#pragma unroll
        for (int c = sub; c < NCH; c += TPR)
          cp_async16(a1 + (c >> 3) * TILES * kTile + sw128_off(r, c & 7), src + c * 8, inside);
      } else {
        constexpr int C1 = CIN / 2;   // channels coming from the upsampled low-res tensor
        const int h = W / 2;
        const float sy = (float)(h - 1) / (float)(W - 1) * (float)(inside ? gy : 0);
        const float sx = (float)(h - 1) / (float)(W - 1) * (float)(inside ? gx : 0);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < h - 1 ? 1 : 0);
        const float wy1 = sy - (float)y0, wy0 = 1.f - wy1, wx1 = sx - (float)x0, wx0 = 1.f - wx1;
        // four tap weights (fp32 products, rounded once to bf16) -> packed bf16x2 FMAs on 2 channels at a time
        const __nv_bfloat162 w00 = __float2bfloat162_rn(wy0 * wx0), w01 = __float2bfloat162_rn(wy0 * wx1),
                             w10 = __float2bfloat162_rn(wy1 * wx0), w11 = __float2bfloat162_rn(wy1 * wx1);
        const __nv_bfloat16* lb = p.low + (size_t)q.b * h * h * C1;
        const __nv_bfloat16* p00 = lb + (size_t)(y0 * h + x0) * C1;
        const __nv_bfloat16* p01 = lb + (size_t)(y0 * h + x1) * C1;
        const __nv_bfloat16* p10 = lb + (size_t)(y1 * h + x0) * C1;
        const __nv_bfloat16* p11 = lb + (size_t)(y1 * h + x1) * C1;
        const __nv_bfloat16* sk = p.in + ((size_t)(q.b * W + (inside ? gy : 0)) * W + (inside ? gx : 0)) * C1;
        // skip-tensor chunks first (asynchronous), then the bilinear chunks with all their loads in flight
#pragma unroll
        for (int c = C1 / 8 + sub; c < NCH; c += TPR)
          cp_async16(a1 + (c >> 3) * TILES * kTile + sw128_off(r, c & 7), sk + (c * 8 - C1), inside);
        constexpr int NBIL = C1 / 8;          // bilinear chunks of this pixel, processed 4 at a time
        constexpr int NB = NBIL < 4 ? NBIL : 4;
        static_assert(NBIL % NB == 0, "bilinear chunk split");
#pragma unroll 1
        for (int c0 = 0; c0 < NBIL; c0 += NB) {
          uint4 ta[NB], tb[NB], tc[NB], td[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int c = c0 + i;
            ta[i] = tb[i] = tc[i] = td[i] = make_uint4(0, 0, 0, 0);
            if (inside) {
              ta[i] = __ldg(reinterpret_cast<const uint4*>(p00 + c * 8));
              tb[i] = __ldg(reinterpret_cast<const uint4*>(p01 + c * 8));
              tc[i] = __ldg(reinterpret_cast<const uint4*>(p10 + c * 8));
              td[i] = __ldg(reinterpret_cast<const uint4*>(p11 + c * 8));
            }
          }
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int c = c0 + i;
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ta[i]);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&tb[i]);
            const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&tc[i]);
            const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&td[i]);
            uint4 o;
            __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              po[j] = __hfma2(w11, pd[j], __hfma2(w10, pc[j], __hfma2(w01, pb[j], __hmul2(w00, pa[j]))));
            sts128(a1 + (c >> 3) * TILES * kTile + sw128_off(r, c & 7), o.x, o.y, o.z, o.w);
          }
        }
      }
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(bar(B_A1FULL + ab));
      T(6);
    };

    for (int pi = 0; pi < n_mine; ++pi) produce_a1(pi);
  } else {
    // =========================================== group A: depthwise 3x3, HID -> A2 ==============================
    // Thread = one 4-channel group (8 B) of one item; 16 groups per pixel, 16 items side by side.
    const int tA = tid - kGroup;
    const int g = tA & 15, slot = tA >> 4;
    const uint32_t chunk = (uint32_t)g >> 1;
    const uint32_t g8 = (g & 1) * 8;
    __nv_bfloat162 wt[9][2], wbias[2];
    auto load_taps = [&](int c) {
      const int ch = c * 64 + 4 * g;
      if constexpr (WS) {   // packed bf16 [CH/8][10][8]: this thread's 4 channels are one half of a 16-byte entry
        const uint32_t tp = sWD + (uint32_t)(ch >> 3) * 160u + (uint32_t)(g & 1) * 8u;
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const uint2 w2 = lds64(tp + t9 * 16);
          wt[t9][0] = *reinterpret_cast<const __nv_bfloat162*>(&w2.x);
          wt[t9][1] = *reinterpret_cast<const __nv_bfloat162*>(&w2.y);
        }
        const uint2 b2 = lds64(tp + 9 * 16);
        wbias[0] = *reinterpret_cast<const __nv_bfloat162*>(&b2.x);
        wbias[1] = *reinterpret_cast<const __nv_bfloat162*>(&b2.y);
      } else {
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) {
        const float4 w4 = lds_f4(sWD + (t9 * CH + ch) * 4);
        wt[t9][0] = __floats2bfloat162_rn(w4.x, w4.y);
        wt[t9][1] = __floats2bfloat162_rn(w4.z, w4.w);
      }
      const float4 b4 = lds_f4(sBD + ch * 4);
      wbias[0] = __floats2bfloat162_rn(b4.x, b4.y);
      wbias[1] = __floats2bfloat162_rn(b4.z, b4.w);
      }
    };
    auto tap = [&](__nv_bfloat162* a, const uint2& v, const __nv_bfloat162* w) {
      a[0] = __hfma2(w[0], *reinterpret_cast<const __nv_bfloat162*>(&v.x), a[0]);
      a[1] = __hfma2(w[1], *reinterpret_cast<const __nv_bfloat162*>(&v.y), a[1]);
    };
    auto finish = [&](__nv_bfloat162* a, uint32_t addr) {
      a[0] = __hmax2(a[0], __hmul2(a[0], kslope));
      a[1] = __hmax2(a[1], __hmul2(a[1], kslope));
      sts64(addr, *reinterpret_cast<uint32_t*>(&a[0]), *reinterpret_cast<uint32_t*>(&a[1]));
    };
    if (NC == 1) load_taps(0);
    int u = 0;
    for (int pi = 0; pi < n_mine; ++pi) {
      for (int c = 0; c < NC; ++c, ++u) {
        const int b = u & 1, k = u >> 1;
        if (NC > 1) load_taps(c);
        T(15);
        mbar_wait(bar(B_HIDFULL + b), k & 1);
        T(12);
        mbar_wait(bar(B_A2FREE + b), (k & 1) ^ 1);
        T(13);
        const uint32_t hid = sHID + b * C::kHidBuf + g8;
        const uint32_t a2 = sA2 + b * C::kA2Buf + g8;
        if constexpr (STRIDE == 1) {
          // item = (output column ox, block of up to 7 output rows): vertical sliding 3x3 window.  Window pixel
          // (hy,hx) is row hy*16+hx of the swizzled tile: its XOR term depends on hx only, so three column bases
          // are computed once per item and rows are reached with immediate offsets (2048 B per window row).
          // Two adjacent output columns per item: 4 window columns feed 2 x 9 taps (4 independent FMA chains).
          constexpr int RB = TOH > 7 ? 7 : 3;         // rows per item: 14 = 2 x 7 (TILES = 2), 6 = 2 x 3 (TILES = 1)
          static_assert(TOH == 2 * RB, "two row blocks per column pair");
          constexpr int NITEM = 14;                   // 7 column pairs x 2 row blocks (<= 16 slots: one pass)
          for (int item = slot; item < NITEM; item += 16) {
            const int ox = (item >> 1) * 2, oy0 = (item & 1) * RB;
            uint32_t cb[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) cb[kx] = hid + oy0 * 2048 + (ox + kx) * 128 + ((chunk ^ ((ox + kx) & 7)) << 4);
            uint2 r0[4], r1[4], r2[4], rn[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              r0[kx] = lds64(cb[kx]);
              r1[kx] = lds64(cb[kx] + 2048);
              r2[kx] = lds64(cb[kx] + 2 * 2048);
            }
            int op = oy0 * 14 + ox;
#pragma unroll
            for (int i = 0; i < RB; ++i) {
              // the window row of the NEXT output row is requested before this row's 36 FMAs, so the shared-memory
              // latency hides behind them
#if FUSED_OPT_DWPF
              if (i + 1 < RB) {
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) rn[kx] = lds64(cb[kx] + (i + 3) * 2048);
              }
#endif
              __nv_bfloat162 a[2] = {wbias[0], wbias[1]}, e[2] = {wbias[0], wbias[1]};
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                tap(a, r0[kx], wt[kx]);
                tap(e, r0[kx + 1], wt[kx]);
                tap(a, r1[kx], wt[3 + kx]);
                tap(e, r1[kx + 1], wt[3 + kx]);
                tap(a, r2[kx], wt[6 + kx]);
                tap(e, r2[kx + 1], wt[6 + kx]);
              }
              finish(a, a2 + op * 128 + ((chunk ^ (op & 7)) << 4));   // A2 tiles are contiguous: row op
              finish(e, a2 + (op + 1) * 128 + ((chunk ^ ((op + 1) & 7)) << 4));
              op += 14;
#if !FUSED_OPT_DWPF
              if (i + 1 < RB) {
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) rn[kx] = lds64(cb[kx] + (i + 3) * 2048);
              }
#endif
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) {
                r0[kx] = r1[kx];
                r1[kx] = r2[kx];
                r2[kx] = rn[kx];
              }
            }
          }
        } else {
          // stride 2: centre of output (oy,ox) is window pixel (2oy+1, 2ox+1)
          for (int item = slot; item < NOUT; item += 16) {
            const int oy = item / 7, ox = item - oy * 7;
            __nv_bfloat162 a[2] = {wbias[0], wbias[1]};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int r = (2 * oy + ky) * 16 + 2 * ox + kx;
                tap(a, lds64(hid + r * 128 + ((chunk ^ (r & 7)) << 4)), wt[ky * 3 + kx]);
              }
            finish(a, a2 + item * 128 + ((chunk ^ (item & 7)) << 4));
          }
        }
        fence_proxy_async();
        mbar_arrive(bar(B_A2FULL + b));
        mbar_arrive(bar(B_HIDFREE + b));
        T(14);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_dealloc(tmem, 512);
}

template <int CIN, int COUT, int STRIDE, bool UPCAT, bool RES, int TILES, int A1BUFS, bool WS = false>
int launch_t(const FusedArgs& a, cudaStream_t st) {
  using C = FCfg<CIN, COUT, STRIDE, TILES, A1BUFS, WS>;
  static unsigned long long attr_devs = 0;   // the attribute is a per-device setting
  auto kfn = fused_ir_kernel<CIN, COUT, STRIDE, UPCAT, RES, TILES, A1BUFS, WS>;
  if (WS && !a.wdp) return (int)cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_devs >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_devs |= 1ull << (dev & 63);
  }
  const int Wo = a.W / STRIDE;
  const int NP = a.batch * ((Wo + C::TOW - 1) / C::TOW) * ((Wo + C::TOH - 1) / C::TOH);
  const int grid = NP < a.num_sms ? NP : a.num_sms;
  return (int)launch_pdl(kfn, dim3(grid), dim3(kThreads), C::kSmem, st, a);
}

}  // namespace

bool fused_ir_supported(int cin, int cout, int stride, bool upcat, bool res) {
  FusedArgs a{};
  a.cin = cin; a.cout = cout; a.stride = stride; a.upcat = upcat; a.res = res; a.batch = 0;
  return launch_fused_ir(a, nullptr) != -1;
}

// returns -1 when the block shape has no fused instantiation, else 0 / cudaError
int launch_fused_ir(const FusedArgs& a, cudaStream_t st) {
#define CASE(CIN_, COUT_, S_, U_, R_, TILES_, A1B_)                                           \
  if (a.cin == CIN_ && a.cout == COUT_ && a.stride == S_ && a.upcat == U_ && a.res == R_) {   \
    if (a.batch <= 0) return 0;                                                               \
    return launch_t<CIN_, COUT_, S_, U_, R_, TILES_, A1B_>(a, st);                            \
  }
  CASE(32, 64, 2, false, false, 2, 2)    // down1.0
  CASE(64, 64, 1, false, true, 2, 1)     // down1.1, up2.1
  CASE(64, 128, 2, false, false, 2, 2)   // down2.0
  CASE(64, 32, 1, true, false, 2, 2)     // up4.0
  CASE(32, 32, 1, false, true, 2, 2)     // up4.1, up3.1
  CASE(32, 64, 1, false, false, 2, 2)    // audio conv1
  CASE(64, 128, 1, false, false, 2, 1)   // audio conv2
  CASE(128, 32, 1, true, false, 1, 2)    // up3.0
#undef CASE
#define CASE_WS(CIN_, COUT_, S_, U_, R_, TILES_, A1B_)                                        \
  if (a.cin == CIN_ && a.cout == COUT_ && a.stride == S_ && a.upcat == U_ && a.res == R_) {   \
    if (a.batch <= 0) return 0;                                                               \
    return launch_t<CIN_, COUT_, S_, U_, R_, TILES_, A1B_, true>(a, st);                      \
  }
  CASE_WS(128, 128, 1, false, true, 1, 2)   // down2.1 (40x40): weights streamed
  CASE_WS(256, 64, 1, true, false, 1, 1)    // up2.0   (40x40): weights streamed
#undef CASE_WS
  return -1;
}

}  // namespace casync
