// tcgen05 GEMM kernel (sm_100a), persistent and warp-specialised.  C[M,N] = epilogue(A[M,K] . W[N,K]^T).
//
//   warps 0-7  A producers   plain row-major A: ONE thread issues a 2-D TMA tensor copy per stage (box 64 x 128,
//                            SWIZZLE_128B, rows beyond M zero-filled by the TMA unit) -- the other producer threads
//                            idle.  Gathered A: 256 threads fill the A stage (128 rows x 64 bf16, SWIZZLE_128B) with
//                            cp.async 16 B chunks (implicit-GEMM 3x3 taps), or values computed on the fly (bilinear x2
//                            upsample of the low-res tensor for the decoder concat, module/unet.py:90-96).  A stage
//                            is published kLag k-blocks after it was issued (cp.async.wait_group + proxy fence +
//                            mbarrier arrive), so several stages of loads stay in flight per thread.
//   warp  8    B loader      one thread: the weight tile (BN rows) is stored in global memory as the swizzled
//                            shared-memory image, so a stage is one bulk-async (TMA engine) copy on the same mbarrier.
//   warp  9    MMA issuer    one thread: 4 x tcgen05.mma (M=128, N=BN, K=16) per k-block into one of two TMEM
//                            accumulators; tcgen05.commit frees the stage / publishes the accumulator.
//   warps 10-17 epilogue     (two per TMEM lane quarter, half of the columns each; residual rows prefetched)
//                            drain TMEM (tcgen05.ld 32x32b) and apply the fused epilogue (folded-BN bias,
//                            LeakyReLU, pre/post residuals, trailing BN) with 16-byte bf16 stores, overlapping the
//                            main loop of the CTA's next tile.
// Grid = min(#tiles, #SMs); tiles are walked N-fastest so CTAs that share an A tile run at the same time.
#include <cstdlib>
#include "gemm_tc.cuh"
#include "gemm_dev.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>

namespace casync {

namespace {

constexpr int kBM = 128;
constexpr int kABytes = kBM * 128;   // 16 KiB: 128 rows x 128 B
constexpr int kProducers = 256;      // 8 producer warps
constexpr int kEpiWarps = 8;          // dedicated epilogue warps (the 8 producer warps join them when A is TMA-loaded)
constexpr int kThreads = kProducers + 2 * 32 + kEpiWarps * 32;
constexpr int kLag = 2;              // A stages in flight per producer thread before publishing

// DWE ("depthwise in the epilogue", BN = 256 only): the GEMM is the first 1x1 conv of an InvertedResidual on a W x W
// image with W*W <= 128 (the 10x10 stages).  Row tiles are FRAME-aligned (tile = the W*W pixels of one frame, the other
// accumulator rows are ignored), so the epilogue holds a frame's hidden tile [W*W, 256 channels]: it goes to shared
// memory as bf16 and the same 8 warps apply the depthwise 3x3 (+ folded-BN bias + LeakyReLU) from there and store the
// block's SECOND hidden tensor -- no depthwise launch, no round trip of the first one (module/unet.py:17-27).
constexpr int kHidTileRows = 100;
template <int BN, bool DWE = false>
struct Cfg {
  static constexpr int kStage = kABytes + BN * 128;
  static constexpr int kHidTile = DWE ? kHidTileRows * BN * 2 : 0;
  static constexpr int kExtra = 1024 /*align*/ + 256 /*barriers*/ + 8192 /*epilogue vectors*/ +
                                16384 /*epilogue store staging: 8 warps x 32 rows x 64 B*/;
  static constexpr int S = DWE ? 3 : (200 * 1024) / kStage > 8 ? 8 : (200 * 1024) / kStage;   // 256:4  128:6  64:8  32:8
  static constexpr int kSmem = S * kStage + kExtra + kHidTile;
  static_assert(kSmem <= 232448, "shared memory overflow");
  static constexpr int kAccCols = BN < 32 ? 32 : BN;
  static constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128
                                   : 2 * kAccCols <= 256 ? 256 : 512;   // tcgen05.alloc wants a power of two
};

// EPI selects the epilogue at compile time, so that its arithmetic is straight-line code (with run-time flags every
// 8-column group was a chain of LDS -> dependent FADD with a branch between every small block: measured 5.5-10 us of a
// 20-30 us launch):  0 bias (+ LeakyReLU)   1 + skip added after the activation (InvertedResidual)
//                    2 + scaled residual before the activation (fc2, b_1)   3 everything by run-time flags (kv, bn7)
enum : int { EPI_PLAIN = 0, EPI_RES_POST = 1, EPI_RES_PRE = 2, EPI_GENERIC = 3 };
template <int BN, bool DWE, int EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const GemmArgs p, const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmC) {
  using C = Cfg<BN, DWE>;
  constexpr bool kGen = EPI == EPI_GENERIC;
  constexpr int S = C::S;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + S * C::kStage;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (S + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  // developer timing: one thread per role accumulates cycles per activity slot
  long long tmark = p.dbg ? clock64() : 0;
  const long long tstart = tmark;
  const bool timed = p.dbg && (tid == 256 || tid == 288 || tid == 320);
  auto T = [&](int slot) {
    if (timed) {
      const long long now = clock64();
      atomicAdd(p.dbg + slot, (unsigned long long)(now - tmark));
      tmark = now;
    }
  };
  const int KB = (p.K + 63) >> 6;
  const int tile_rows = DWE ? p.dw_w * p.dw_w : kBM;   // DWE: one frame per row tile
  const int NT = p.N / BN, MT = (p.M + tile_rows - 1) / tile_rows;
  const int n_tiles = NT * MT;
  // epilogue column groups (4 warps each, one per TMEM lane quarter).  Letting the idle producer warps of the
  // TMA-loaded mode drain too (16 epilogue warps) was measured SLOWER: a warp's chunk is a serial latency chain
  // (tcgen05.ld -> bias -> activation -> staging -> store), and the extra staging smem costs a pipeline stage.
  const int ncg = BN >= 64 ? 2 : 1;
  const int epi_n = ncg * 128;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), p.amode == A_PLAIN ? 2 : kProducers + 1);   // PLAIN: the TMA lane + the B loader
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), epi_n);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // weights are constant: the B loader (and the MMA issuer behind it) start at once; every role that reads or
  // writes an activation buffer first waits for the previous kernel of the stream
  T(10);
  if (warp < 8 || warp >= 10) pdl_wait();
  if (tid == 320) T(11);

  // ======================================= epilogue (warps 10-17) =================================================
  auto run_epilogue = [&](const int ew) {
    // ======================================= epilogue warps ==========================================================
    const int lg = warp & 3;                 // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                  // column group of this warp
    if (cg >= ncg) return;
    const int cols_per = BN / ncg;
    // per-tile epilogue vectors (bias, residual scale, trailing BN) are staged in shared memory while the tile's
    // main loop runs, double-buffered across tiles: [buffer][bias | rscale | post_scale | post_shift][256]
    float* const evec = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));
    const int et = ew * 32 + lane;
    // Output rows go through a warp-private staging slab (32 rows x 64 B, 16-byte chunks XOR-swizzled by row pair) and
    // leave as 8 rows x 64 contiguous bytes per store instruction: a thread owns one accumulator ROW, so storing
    // straight from registers writes 32 half-filled sectors per instruction (measured: ~11k cycles per 128x256 tile).
    // The slab is a SWIZZLE_64B TMA box (32 columns x 32 rows): one elected lane ships it with a tensor store, which also
    // clips the rows beyond M.  The XOR term is a function of the ABSOLUTE shared-memory address (bits 7-8).
    const uint32_t stg_a = bar_base + 256 + 8192 + ew * 2048;
    uint8_t* const stg = smem_raw + (stg_a - smem_u32(smem_raw));
    const uint32_t sx = ((stg_a + lane * 64) >> 7) & 3;   // swizzle term of this thread's row
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int ab = t & 1;
      const int n0 = (tile % NT) * BN, m0 = (tile / NT) * tile_rows;
      float* const ev = evec + (t & 1) * 1024;
      if (et < BN) {
        ev[et] = __ldg(p.bias + n0 + et);
        if (p.res_pre) ev[256 + et] = __ldg(p.rscale + n0 + et);
        if (p.post_scale) {
          ev[512 + et] = __ldg(p.post_scale + n0 + et);
          ev[768 + et] = __ldg(p.post_shift + n0 + et);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(epi_n) : "memory");
      T(7);
      const int m = m0 + lg * 32 + lane;
      const bool row_ok = m < p.M;
      const int cbeg = cg * cols_per, cend = cbeg + cols_per;
      if constexpr (DWE) {
        // ---- pass 1: hidden tile = leaky(acc + b1) as bf16 -> shared memory [pixel][BN channels]; the 16-byte chunks
        //      of a row are XOR-swizzled with the row index (row pitch 512 B: conflict-free 512-byte warp accesses)
        const int r = lg * 32 + lane;                  // accumulator row = pixel of this tile's frame
        const bool px_ok = r < tile_rows && row_ok;
        uint8_t* const hid = smem_raw + (bar_base + 256 + 8192 + 16384 - smem_u32(smem_raw));
        // taps + bias of the 8 channels this thread handles in pass 2 (constants: requested before the accumulator wait)
        const int dc = et & (BN / 8 - 1), dr0 = et / (BN / 8);
        uint4 wt[9], wb;
        {
          const uint4* tp = reinterpret_cast<const uint4*>(p.dwp) + (size_t)((n0 >> 3) + dc) * 10;
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) wt[t9] = __ldg(tp + t9);
          wb = __ldg(tp + 9);
        }
        const __nv_bfloat162 slope2 = __floats2bfloat162_rn(p.leaky ? kLeaky : 1.0f, p.leaky ? kLeaky : 1.0f);
        mbar_wait(acc_full(ab), (t >> 1) & 1);
        T(8);
        tc_fence_after();
        const uint32_t trow = tmem + ab * C::kAccCols + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          uint32_t acc[32];
          tmem_ld32(trow + c0, acc);
          tmem_ld_wait32(acc);
          if (px_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float v[8];
              const float4 b0 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g);
              const float4 b1 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(acc[8 * g + q]) + bb[q];
              const uint32_t ci = (uint32_t)(c0 >> 3) + g;
              *reinterpret_cast<uint4*>(hid + r * (BN * 2) + ((ci ^ (uint32_t)(r & 7)) << 4)) =
                  make_uint4(pack_leaky(v[0], v[1], slope2), pack_leaky(v[2], v[3], slope2), pack_leaky(v[4], v[5], slope2),
                             pack_leaky(v[6], v[7], slope2));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(acc_empty(ab));                    // the accumulator is free: the next tile's main loop may proceed
        asm volatile("bar.sync 1, %0;" ::"r"(epi_n) : "memory");   // hidden tile complete
        T(13);
        // ---- pass 2: depthwise 3x3 + folded-BN bias + LeakyReLU from the hidden tile (same arithmetic and order as
        //      dw3x3_kernel: taps outside the image contribute w * 0), 16-byte stores of whole 512-byte row pieces
        const int Wd = p.dw_w;
        const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);
#pragma unroll 1
        for (int px = dr0; px < tile_rows; px += kEpiWarps * 32 / (BN / 8)) {
          const int y = px / Wd, x = px - y * Wd;
          __nv_bfloat162 a4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) a4[q] = reinterpret_cast<const __nv_bfloat162*>(&wb)[q];
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            const int yy = y + t9 / 3 - 1, xx = x + t9 % 3 - 1;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (yy >= 0 && yy < Wd && xx >= 0 && xx < Wd) {
              const int rr = yy * Wd + xx;
              v = *reinterpret_cast<const uint4*>(hid + rr * (BN * 2) + (((uint32_t)dc ^ (uint32_t)(rr & 7)) << 4));
            }
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
            const __nv_bfloat162* pw = reinterpret_cast<const __nv_bfloat162*>(&wt[t9]);
#pragma unroll
            for (int q = 0; q < 4; ++q) a4[q] = __hfma2(pw[q], pv[q], a4[q]);
          }
          uint4 o;
          __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) po[q] = __hmax2(a4[q], __hmul2(a4[q], kslope));
          if (m0 + px < p.M) *reinterpret_cast<uint4*>(p.C + (size_t)(m0 + px) * p.ldc + n0 + dc * 8) = o;
        }
        T(15);
        // (the bar.sync at the top of the next tile orders these reads before the next pass 1 overwrites the tile)
        continue;
      }
      // residual rows (a launch uses res_pre or res_post, never both): the first 32-column chunk is requested BEFORE
      // waiting for the accumulator, so its latency hides behind the tile's main loop; later chunks are prefetched
      // one chunk ahead, behind the TMEM load + arithmetic of the current chunk
      const bool has_pre = kGen ? p.res_pre != nullptr : EPI == EPI_RES_PRE;
      const bool has_post = kGen ? p.res_post != nullptr : EPI == EPI_RES_POST;
      const bool has_ps = kGen ? p.post_scale != nullptr : false;
      const bool has_vt = kGen ? (p.vt && n0 >= p.vt_col0) : false;
      const float slope = p.leaky ? kLeaky : 1.0f;   // max(v, 1 * v) == v: LeakyReLU on / off without a branch
      const __nv_bfloat162 slope2 = __floats2bfloat162_rn(slope, slope);
      const __nv_bfloat16* rsrc = has_pre ? p.res_pre + (size_t)m * p.ld_rpre : has_post ? p.res_post + (size_t)m * p.ld_rpost : nullptr;
      uint4 rnext[4];
      auto fetch_res = [&](int c0) {
        if constexpr (EPI != EPI_PLAIN) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            rnext[g] = (rsrc && row_ok) ? *reinterpret_cast<const uint4*>(rsrc + n0 + c0 + 8 * g) : make_uint4(0, 0, 0, 0);
        }
      };
      if (cbeg < cend) fetch_res(cbeg);
      mbar_wait(acc_full(ab), (t >> 1) & 1);
      T(8);
      tc_fence_after();
      const uint32_t trow = tmem + ab * C::kAccCols + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(trow + c0, acc);
        uint4 rcur[4];
        if constexpr (EPI != EPI_PLAIN) {
#pragma unroll
          for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
          if (c0 + 32 < cend) fetch_res(c0 + 32);
        }
        tmem_ld_wait32(acc);
        T(13);
        if (!has_vt) {   // the previous chunk's tensor store has read the slab (its issuer waits, the warp follows)
          if (elect_one()) tma_store_wait_read();
          __syncwarp();
        }
        if (row_ok) {
          const int n = n0 + c0;
          // all arithmetic first, the four staging stores after it: a shared-memory load (bias / scale vectors) cannot
          // be scheduled above an earlier shared-memory store that might alias, so store-per-group serialised the groups
          // (only where registers allow: the scaled-residual and generic variants keep one store per group -- measured)
          constexpr bool kDefer = EPI == EPI_PLAIN || EPI == EPI_RES_POST;
          uint4 og[kDefer ? 4 : 1];
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 8 columns per group -> one 16 B store
            float v[8];
            const float4 b0 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g);
            const float4 b1 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(acc[8 * g + q]) + bb[q];
            if (has_pre) {
              const uint32_t* pr = &rcur[g].x;
              const float4 s0 = *reinterpret_cast<const float4*>(ev + 256 + c0 + 8 * g);
              const float4 s1 = *reinterpret_cast<const float4*>(ev + 256 + c0 + 8 * g + 4);
              const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] += ss[2 * q] * bf16_lo(pr[q]);
                v[2 * q + 1] += ss[2 * q + 1] * bf16_hi(pr[q]);
              }
            }
            if constexpr (kGen) {
              if (p.leaky) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], kLeaky * v[q]);
              }
            } else if constexpr (EPI != EPI_PLAIN) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], slope * v[q]);
            }   // EPI_PLAIN: LeakyReLU on the packed bf16 pairs below (half the instructions)
            if (has_post) {
              const uint32_t* pr = &rcur[g].x;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] += bf16_lo(pr[q]);
                v[2 * q + 1] += bf16_hi(pr[q]);
              }
            }
            if (has_ps) {
              const float4 s0 = *reinterpret_cast<const float4*>(ev + 512 + c0 + 8 * g);
              const float4 s1 = *reinterpret_cast<const float4*>(ev + 512 + c0 + 8 * g + 4);
              const float4 t0 = *reinterpret_cast<const float4*>(ev + 768 + c0 + 8 * g);
              const float4 t1 = *reinterpret_cast<const float4*>(ev + 768 + c0 + 8 * g + 4);
              const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                v[q] = ss[q] * v[q] + tt[q];
                v[q] = fmaxf(v[q], kLeaky * v[q]);
              }
            }
            if (has_vt) {   // transposed store of a value projection (see GemmArgs::vt)
              const int frame = m / 100, key = m - frame * 100;
              const int colv = n + 8 * g - p.vt_col0;   // j * 512 + channel
              __nv_bfloat16* dst = p.vt + ((size_t)frame * 2048 + colv) * 128 + key;
#pragma unroll
              for (int q = 0; q < 8; ++q) dst[q * 128] = __float2bfloat16_rn(v[q]);
            } else {
              uint4& o = og[kDefer ? g : 0];
              if constexpr (EPI == EPI_PLAIN) {
                o.x = pack_leaky(v[0], v[1], slope2);
                o.y = pack_leaky(v[2], v[3], slope2);
                o.z = pack_leaky(v[4], v[5], slope2);
                o.w = pack_leaky(v[6], v[7], slope2);
              } else {
                o.x = pack_bf16(v[0], v[1]);
                o.y = pack_bf16(v[2], v[3]);
                o.z = pack_bf16(v[4], v[5]);
                o.w = pack_bf16(v[6], v[7]);
              }
              if constexpr (!kDefer) *reinterpret_cast<uint4*>(stg + lane * 64 + (((uint32_t)g ^ sx) << 4)) = o;
            }
          }
          if constexpr (kDefer) {
#pragma unroll
            for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stg + lane * 64 + (((uint32_t)g ^ sx) << 4)) = og[g];
          }
        }
        T(14);
        if (!has_vt) {
          fence_proxy_async();   // generic-proxy slab writes -> async proxy
          __syncwarp();
          if (elect_one()) tma_store_2d(&tmC, stg_a, n0 + c0, m0 + lg * 32);   // converged warp: same lane every time
        }
        T(15);
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));
      T(9);
    }
    __syncwarp();
    if (elect_one()) tma_store_wait_read();   // shared memory must outlive the last tensor store
  };

  if (warp < 8 && p.amode == A_PLAIN) {
    // ======================================= A via TMA (one elected lane of warp 0) =================================
    if (warp == 0) {
      int j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / NT) * tile_rows;
        for (int kb = 0; kb < KB; ++kb, ++j) {
          const int s = j % S;
          mbar_wait(empty(s), ((j / S) & 1) ^ 1);
          if (elect_one()) {   // one 2-D TMA tensor copy (box 64 x 128, rows beyond M zero-filled)
            mbar_arrive_expect_tx(full(s), kABytes);
            tma_load_2d(base + s * C::kStage, &tmA, kb * 64, m0, full(s));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    // ======================================= A producers (gathered operand) ========================================
    // PLAIN / CONV3X3: thread owns chunk (tid & 7) of rows (tid >> 3) + 32 i, i < 4 (a warp copies four
    // full 128 B rows per instruction).  UPCAT: two threads per row (4 chunks each; 4 taps + weights per row).
    int j = 0;           // k-blocks issued by this thread over all tiles (stage = j % S)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / NT) * kBM;
      int conv_pix[4], conv_yx[4];
      RowCoord rc{};
      if (p.amode == A_CONV3X3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int m = m0 + (tid >> 3) + 32 * i;
          int mm = m < p.M ? m : 0;
          int ox = mm % p.Wout, t = mm / p.Wout;
          int oy = t % p.Hout, b = t / p.Hout;
          int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
          conv_pix[i] = (b * p.Hin + iy0) * p.Win + ix0;
          conv_yx[i] = m < p.M ? (((iy0 + 64) << 16) | (ix0 + 64)) : -1;
        }
      } else if (p.amode == A_UPCAT) {
        int m = m0 + (tid >> 1);
        int mm = m < p.M ? m : 0;
        int x = mm % p.Wout, t = mm / p.Wout;
        int y = t % p.Hout, b = t / p.Hout;
        // align_corners=True source coordinates (ATen area_pixel_compute_scale / upsample_bilinear2d)
        float sy = (float)(p.Hin - 1) / (float)(p.Hout - 1) * (float)y;
        float sx = (float)(p.Win - 1) / (float)(p.Wout - 1) * (float)x;
        int y0 = (int)sy, x0 = (int)sx;
        int y1 = y0 + (y0 < p.Hin - 1 ? 1 : 0), x1 = x0 + (x0 < p.Win - 1 ? 1 : 0);
        rc.wy1 = sy - (float)y0;
        rc.wy0 = 1.f - rc.wy1;
        rc.wx1 = sx - (float)x0;
        rc.wx0 = 1.f - rc.wx1;
        const __nv_bfloat16* fb = p.A + (size_t)b * p.Hin * p.Win * p.Cin;
        rc.p00 = fb + (size_t)(y0 * p.Win + x0) * p.Cin;
        rc.p01 = fb + (size_t)(y0 * p.Win + x1) * p.Cin;
        rc.p10 = fb + (size_t)(y1 * p.Win + x0) * p.Cin;
        rc.p11 = fb + (size_t)(y1 * p.Win + x1) * p.Cin;
      }
      for (int kb = 0; kb < KB; ++kb, ++j) {
        const int s = j % S;
        mbar_wait(empty(s), ((j / S) & 1) ^ 1);
        const uint32_t a_s = base + s * C::kStage;
        if (p.amode == A_UPCAT) {
          const int r = tid >> 1, m = m0 + r;
          const int c2 = p.K - p.Cin;
#pragma unroll
          for (int c = (tid & 1) * 4; c < (tid & 1) * 4 + 4; ++c) {
            const int k = (kb * 8 + c) * 8;
            const uint32_t dst = a_s + sw128_off(r, c);
            if (m < p.M && k < p.Cin) {
              uint4 a = __ldg(reinterpret_cast<const uint4*>(rc.p00 + k));
              uint4 b = __ldg(reinterpret_cast<const uint4*>(rc.p01 + k));
              uint4 cc = __ldg(reinterpret_cast<const uint4*>(rc.p10 + k));
              uint4 d = __ldg(reinterpret_cast<const uint4*>(rc.p11 + k));
              uint4 o = lerp8(a, b, cc, d, rc);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z),
                           "r"(o.w));
            } else {
              bool valid = m < p.M && k < p.K;
              const __nv_bfloat16* src = valid ? p.A2 + (size_t)m * c2 + (k - p.Cin) : p.A2;
              cp_async16(dst, src, valid);
            }
          }
        } else {
          const int c = tid & 7;
          const int k = (kb * 8 + c) * 8;
          if (p.amode == A_PLAIN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = (tid >> 3) + 32 * i, m = m0 + r;
              bool valid = m < p.M && k < p.K;
              const __nv_bfloat16* src = valid ? p.A + (size_t)m * p.lda + k : p.A;
              cp_async16(a_s + sw128_off(r, c), src, valid);
            }
          } else {  // implicit GEMM over the 9 taps of a dense 3x3 conv: k = tap*Cin + ci
            const int tap = k / p.Cin, ci = k - tap * p.Cin;
            const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = (tid >> 3) + 32 * i;
              const int iy = (conv_yx[i] >> 16) - 64 + ky, ix = (conv_yx[i] & 0xFFFF) - 64 + kx;
              bool valid = conv_yx[i] >= 0 && k < p.K && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
              const __nv_bfloat16* src = valid ? p.A + (size_t)(conv_pix[i] + ky * p.Win + kx) * p.Cin + ci : p.A;
              cp_async16(a_s + sw128_off(r, c), src, valid);
            }
          }
        }
        cp_async_commit();
        if (j >= kLag) {   // publish the stage issued kLag k-blocks ago
          cp_async_wait<kLag>();
          fence_proxy_async();
          mbar_arrive(full((j - kLag) % S));
        }
      }
    }
    // drain: publish the last kLag stages
    cp_async_wait<0>();
    fence_proxy_async();
    for (int q = (j > kLag ? j - kLag : 0); q < j; ++q) mbar_arrive(full(q % S));
  } else if (warp == 8) {
    // ======================================= B loader (one elected lane) =============================================
    {
      int j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = (tile % NT) * BN;
        for (int kb = 0; kb < KB; ++kb, ++j) {
          const int s = j % S;
          mbar_wait(empty(s), ((j / S) & 1) ^ 1);
          T(2);
          if (elect_one()) {
            mbar_arrive_expect_tx(full(s), BN * 128);
            bulk_g2s(base + s * C::kStage + kABytes, p.W + ((size_t)kb * p.N + n0) * 128, BN * 128, full(s));
          }
          __syncwarp();
          T(3);
        }
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer (one elected lane of a converged warp) ======================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int j = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int ab = t & 1;
        mbar_wait(acc_empty(ab), ((t >> 1) & 1) ^ 1);
        T(4);
        tc_fence_after();
        const uint32_t d = tmem + ab * C::kAccCols;
        for (int kb = 0; kb < KB; ++kb, ++j) {
          const int s = j % S;
          mbar_wait(full(s), (j / S) & 1);
          T(5);
          tc_fence_after();
          const uint32_t a_s = base + s * C::kStage;
          const uint64_t adesc = umma_desc_sw128(a_s), bdesc = umma_desc_sw128(a_s + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kb | ks) != 0);
            umma_commit(empty(s));
            if (kb == KB - 1) umma_commit(acc_full(ab));
          }
          __syncwarp();
          T(6);
        }
      }
    }
  } else {
    run_epilogue(warp - 10);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, C::kTmemCols);
  if (p.dbg && tid == 0) atomicAdd(p.dbg + 12, (unsigned long long)(clock64() - tstart));
}

int g_num_sms = 148;
thread_local int g_cost_cap = 0;   // gemm_set_cost_cap

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

int gemm_encode_map(void* tm_out, const __nv_bfloat16* A, int rows, int cols, int ld, int box_rows, bool swizzle128) {
  // [rows, cols] bf16 row-major with pitch ld: box = 64 columns (128 B) x box_rows rows; rows outside are zero-filled
  if (!g_encode || (ld & 7) || ((uintptr_t)A & 15) || box_rows < 1 || box_rows > 256) return (int)cudaErrorInvalidValue;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  if (g_encode(reinterpret_cast<CUtensorMap*>(tm_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(A),
               gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  return 0;
}

static int gemm_encode_store_map(void* tm_out, const __nv_bfloat16* Cp, int rows, int cols, int ld) {
  if (!g_encode || (ld & 7) || ((uintptr_t)Cp & 15)) return (int)cudaErrorInvalidValue;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  if (g_encode(reinterpret_cast<CUtensorMap*>(tm_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(Cp),
               gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  return 0;
}

int gemm_num_sms() { return g_num_sms; }
void gemm_set_cost_cap(int ctas) { g_cost_cap = ctas; }

namespace {

template <int BN, bool DWE = false, int EPI = EPI_GENERIC>
int launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  const int tile_rows = DWE ? a.dw_w * a.dw_w : kBM;
  const int tiles = (a.N / BN) * ((a.M + tile_rows - 1) / tile_rows);
  const int cap = a.max_ctas > 0 && a.max_ctas < g_num_sms ? a.max_ctas : g_num_sms;
  const int grid = tiles < cap ? tiles : cap;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (a.amode == A_PLAIN) {
    // A[M, K] bf16 row-major with pitch lda: box = 64 columns (128 B, one swizzle span) x 128 rows
    const int e = gemm_encode_map(&tm, a.A, a.M, a.K, a.lda, kBM, true);
    if (e) return e;
  }
  // C[M, N] bf16 row-major with pitch ldc: box = 32 columns (64 B) x 32 rows, SWIZZLE_64B (the epilogue's staging slab)
  alignas(64) CUtensorMap tc;
  memset(&tc, 0, sizeof tc);
  if (!DWE) {
    const int e = gemm_encode_store_map(&tc, a.C, a.M, a.N, a.ldc);
    if (e) return e;
  }
  return (int)launch_pdl(gemm_tc_kernel<BN, DWE, EPI>, dim3(grid), dim3(kThreads), Cfg<BN, DWE>::kSmem, stream, a, tm, tc);
}

template <int BN, bool DWE, int EPI>
int set_attr1() {
  return (int)cudaFuncSetAttribute(gemm_tc_kernel<BN, DWE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg<BN, DWE>::kSmem);
}
template <int BN>
int set_attr() {
  return set_attr1<BN, false, EPI_PLAIN>() | set_attr1<BN, false, EPI_RES_POST>() | set_attr1<BN, false, EPI_RES_PRE>() |
         set_attr1<BN, false, EPI_GENERIC>();
}
// the epilogue variant a launch needs (see gemm_tc_kernel)
int epi_of(const GemmArgs& a) {
  if (a.post_scale || a.vt || (a.res_pre && a.res_post)) return EPI_GENERIC;
  return a.res_pre ? EPI_RES_PRE : a.res_post ? EPI_RES_POST : EPI_PLAIN;
}
template <int BN>
int launch_bn(const GemmArgs& a, cudaStream_t stream) {
  switch (epi_of(a)) {
    case EPI_PLAIN: return launch_cfg<BN, false, EPI_PLAIN>(a, stream);
    case EPI_RES_POST: return launch_cfg<BN, false, EPI_RES_POST>(a, stream);
    case EPI_RES_PRE: return launch_cfg<BN, false, EPI_RES_PRE>(a, stream);
    default: return launch_cfg<BN, false, EPI_GENERIC>(a, stream);
  }
}

// work-per-SM proxy: waves of tiles x (per-tile cost ~ A bytes + B bytes per k-block, the L2->smem traffic)
double tile_cost(long mt, int N, int bn, int cap) {
  if (N % bn) return 1e30;
  const long tiles = mt * (N / bn);
  const long waves = (tiles + cap - 1) / cap;
  return (double)waves * (kABytes + bn * 128 + 6000 /*fixed per-k-block/tile overhead*/);
}

}  // namespace

int gemm_init() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    g_num_sms = sms;
  int e = 0;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !fn)
      return (int)cudaErrorNotSupported;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  e |= set_attr<32>();
  e |= set_attr<64>();
  e |= set_attr<128>();
  e |= set_attr<192>();
  e |= set_attr<256>();
  e |= set_attr1<256, true, EPI_PLAIN>();
  return e;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  // partial k-blocks: the packed weights are zero-padded to 64, TMA zero-fills columns >= K, gathers test k < K
  if (a.N % 32 != 0 || a.K % 8 != 0) return (int)cudaErrorInvalidValue;
  if (a.dw_epi) {   // depthwise in the epilogue: frame-aligned tiles, BN = 256 (see Cfg)
    const int px = a.dw_w * a.dw_w;
    if (a.amode != A_PLAIN || a.N % 256 != 0 || px < 1 || px > kHidTileRows || a.M % px != 0 || !a.dwp || a.ldc != a.N ||
        a.res_pre || a.res_post || a.post_scale || a.vt || ((uintptr_t)a.dwp & 15))
      return (int)cudaErrorInvalidValue;
    return launch_cfg<256, true, EPI_PLAIN>(a, stream);
  }
  const long mt = (a.M + kBM - 1) / kBM;
  const int cap = a.max_ctas > 0 && a.max_ctas < g_num_sms ? a.max_ctas : g_num_sms;
  // tile shape: while two lanes share the GPU (plan.cu, split batches) a launch gets about half of the SMs, and the wave
  // count is computed for that share (measured: 1.798 -> 1.789 ms at batch 64, 1.082 -> 1.059 at 32, 6.567 -> 6.495 at 256;
  // for an unsplit batch 8 the same assumption costs 4 %, so it is tied to the split)
  const int ccap = g_cost_cap > 0 && g_cost_cap < cap ? g_cost_cap : cap;
  int best = 32;
  double bc = tile_cost(mt, a.N, 32, ccap);
  static const int bn_max = getenv("CASYNC_GEMM_BNMAX") ? atoi(getenv("CASYNC_GEMM_BNMAX")) : 256;   // developer A/B
  for (int bn : {64, 128, 192, 256}) {   // 192: N = 576 (p_1 + q) -> 3 column tiles instead of 9
    if (bn > bn_max) continue;
    if (a.vt && a.vt_col0 % bn) continue;   // a column tile must not straddle the row-major | transposed boundary
    const double c = tile_cost(mt, a.N, bn, ccap);
    if (c <= bc) {
      bc = c;
      best = bn;
    }
  }
  switch (best) {
    case 256: return launch_bn<256>(a, stream);
    case 192: return launch_bn<192>(a, stream);
    case 128: return launch_bn<128>(a, stream);
    case 64: return launch_bn<64>(a, stream);
    default: return launch_bn<32>(a, stream);
  }
}

}  // namespace casync
