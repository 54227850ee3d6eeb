// Layer-program kernel (sm_100a): one persistent, warp-specialised launch executes the work items of MANY layers.
//
//   item = (layer, row tile mt of 128 rows, column tile nt); items are numbered layer by layer and CTA c executes items
//   c, c + grid, c + 2 grid, ... in order.  Every role walks the same item sequence on its own:
//
//   warps 0-7   GEMM, plain A:     warp 0: waits for the producer row tiles (acquire loads of their completion
//                                  counters), then one 2-D TMA tensor copy per k-block (box 64 x 128, SWIZZLE_128B)
//               GEMM, gathered A:  256 threads fill the A stage with cp.async 16 B chunks (implicit-GEMM 3x3 taps) or
//                                  with the bilinear x2 upsample + skip concat of the decoder computed on the fly
//               depthwise item:    256 threads: slab of hidden rows [m0 - W - 1, m0 + 128 + W + 1) x 64 channels per
//                                  TMA copy (double-buffered), 9 LDS.128 + packed bf16x2 FMAs per (row, 8 channels),
//                                  16-byte stores; stride 2: taps read straight from L2
//   warp  8     B loader           weight tiles (pre-swizzled shared-memory image in global memory): one bulk copy per
//                                  k-block; weights are constants, so it runs ahead of the dependencies
//   warp  9     MMA issuer         4 x tcgen05.mma (M 128, N = BN of the layer, K 16) per k-block into one of two TMEM
//                                  accumulators (256 columns each)
//   warps 10-17 epilogue           TMEM -> bias / LeakyReLU / residuals / trailing BN -> bf16 rows through swizzled
//                                  staging slabs -> global; then fence + one release-add on the row tile's counter
//
// Dependencies replace kernel boundaries: layer L+1's tile starts when the row tiles it reads are complete, so the
// epilogue of a tile overlaps the main loop of the CTA's next item (usually of ANOTHER layer), layer tails overlap
// the next layer's head, and there is no launch / prologue / wave-quantisation cost per layer.  Deadlock-free: items
// only depend on lower-numbered items, every CTA is resident (grid <= #SMs, 1 CTA/SM) and executes in order.
// Buffers written inside one program are written exactly once (the host allocates the hidden tensors of every
// InvertedResidual from an arena), so there are no WAR hazards between unordered items.
#include "chain.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gemm_dev.cuh"

namespace casync {

namespace {

constexpr int kBM = 128;
constexpr int kABytes = kBM * 128;
constexpr int kProducers = 256;
constexpr int kThreads = kProducers + 2 * 32 + 8 * 32;
constexpr int kLag = 2;
constexpr int kS = 3;                            // pipeline stages of 48 KiB (A 16 KiB + B up to 256 rows x 128 B)
constexpr int kStage = kABytes + 256 * 128;
constexpr int kSlabRows = 210;                   // depthwise slab: 130 + 2 W rows of 128 B (W <= 40) ...
constexpr int kSlabTaps = kSlabRows * 128;       // ... followed by the slice's taps + bias: 8 chunks x 10 x 16 B (bf16)
constexpr int kSlab = kSlabTaps + 1280;
constexpr int kOffBar = kS * kStage;
constexpr int kOffEvec = kOffBar + 256;
constexpr int kOffStg = kOffEvec + 8192;
constexpr int kOffSlab = kOffStg + 16384;
constexpr int kOffTab = kOffSlab + 2 * kSlab;
struct LayerS {   // what the roles need per item, copied to shared memory once per launch
  int item_end, item0, NT, KB, cnt0, need, out_rows;
  short kind, amode, BN, pad_;
  short d_mode[2], d_layer[2];
  int d_a[2], d_b[2];
};
constexpr int kSmem = 1024 + kOffTab + (int)sizeof(LayerS) * kChainMaxLayers;
static_assert(kSmem <= 232448, "shared memory overflow");
constexpr int kAccCols = 256;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// one thread: block until every producer row tile this item reads is complete
__device__ __forceinline__ void wait_deps(const LayerS* tab, const LayerS& L, int m0, const unsigned* cnt) {
  const int m1 = min(m0 + kBM - 1, L.out_rows - 1);
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int mode = L.d_mode[d];
    if (mode == 0) continue;
    const LayerS& P = tab[L.d_layer[d]];
    const int a = L.d_a[d], b = L.d_b[d];
    int r_lo, r_hi;
    if (mode == 1) {
      r_lo = m0 - a;
      r_hi = m1 + a;
    } else {
      r_lo = (m0 / a) * b;
      r_hi = (m1 / a + 1) * b - 1;
    }
    r_lo = max(r_lo, 0);
    r_hi = min(r_hi, P.out_rows - 1);
    const unsigned need = (unsigned)P.need;
    const unsigned* c = cnt + P.cnt0;
    for (int t = r_hi >> 7; t >= (r_lo >> 7); --t)   // the last tile is the most likely to be late
      while (ld_acquire_u32(c + t) < need) __nanosleep(32);
  }
  // the data was written through the generic proxy (st.global of other CTAs); TMA reads it through the async proxy
  asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const ChainLayer* __restrict__ layers, int n_layers,
                                                            int total_items, unsigned* __restrict__ cnt,
                                                            unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + kOffBar;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (kS + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * kS + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * kS + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kS + 4);
  auto slab_full = [&](int i) { return bar_base + 8u * (2 * kS + 5 + i); };
  // a_go(s): "stage s is free for the k-block of a GATHERED A operand".  The workers skip plain-GEMM items without
  // waiting, so they may be several ring phases ahead when they reach a gathered item -- a parity wait on empty(s)
  // would alias to an older phase.  The loader walks the ring in order and hands each such stage over.
  auto a_go = [&](int s) { return bar_base + 8u * (2 * kS + 7 + s); };
  LayerS* const tab = reinterpret_cast<LayerS*>(gbase + kOffTab);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  // developer timing (CASYNC_CHAIN_DBG=1): one thread per role accumulates cycles per activity slot
  long long tmark = dbg ? clock64() : 0;
  const bool timed = dbg && (tid == 0 || tid == 256 || tid == 288 || tid == 320);
  auto T = [&](int slot) {
    if (timed) {
      const long long now = clock64();
      atomicAdd(dbg + slot, (unsigned long long)(now - tmark));
      tmark = now;
    }
  };
  for (int i = tid; i < n_layers; i += kThreads) {   // descriptors are written by the host before the launch
    const ChainLayer* L = layers + i;
    LayerS s;
    s.item_end = L->item0 + L->items;
    s.item0 = L->item0;
    s.kind = L->kind;
    s.amode = L->g.amode;
    s.BN = L->BN;
    s.NT = L->NT;
    s.KB = L->g.K >> 6;
    s.cnt0 = L->cnt0;
    s.need = L->need;
    s.out_rows = L->out_rows;
    for (int d = 0; d < 2; ++d) {
      s.d_mode[d] = L->dep[d].mode;
      s.d_layer[d] = L->dep[d].layer;
      s.d_a[d] = L->dep[d].a;
      s.d_b[d] = L->dep[d].b;
    }
    tab[i] = s;
  }
  if (tid == 0) {
    for (int s = 0; s < kS; ++s) {
      mbar_init(full(s), 2);     // A (loader's TMA copy, or thread 0 of the gathering workers) + B (loader)
      mbar_init(empty(s), 1);
      mbar_init(a_go(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 256);
      mbar_init(slab_full(a), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // every role that reads or writes an activation buffer first waits for the previous kernel of the stream
  if (warp != 9) pdl_wait();

  if (warp < 8) {
    // =========================== workers (warps 0-7): depthwise items, gathered A operands ========================
    int j = 0;       // k-blocks of GEMM items so far (stage = j % kS)
    int su = 0;      // slab loads so far (buffer = su & 1)
    int li = 0;
    uint32_t gph = 0;   // bit s: parity of the next a_go(s) phase
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      while (it >= tab[li].item_end) ++li;
      const LayerS& S = tab[li];
      if (S.kind == CK_GEMM && S.amode == A_PLAIN) {   // the loader warp feeds this item
        j += S.KB;
        continue;
      }
      const ChainLayer* __restrict__ L = layers + li;
      const int local = it - S.item0;
      const int NT = S.NT;
      const int mt = local / NT, nt = local - mt * NT;
      const int m0 = mt * kBM;
      T(15);
      __syncwarp();   // elect.sync below names the full warp: reconverge after the divergent tail of the last item
      if (warp == 0) {
        if (elect_one()) wait_deps(tab, S, m0, cnt);
        __syncwarp();
      }
      bar_sync(2, kProducers);
      T(10);
      if (S.kind == CK_DW) {
        // ----------------------------------- depthwise 3x3 + bias + LeakyReLU ------------------------------------
        const int K = L->g.K, W = L->g.Win, H = L->g.Hin, M = L->g.M;
        const int c = tid & 7;
        const int s0 = nt * 4, ns = min(4, (K >> 6) - s0);
        const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);
        const uint8_t* wdp = L->g.W;   // taps + bias as bf16, [K/8][10][8]: 1280 contiguous bytes per 64-channel slice
        __nv_bfloat16* const outp = L->g.C;
        if (L->g.stride == 1) {
          const uint32_t slab_bytes = (uint32_t)(130 + 2 * W) * 128u;
          uint32_t vmask[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = m0 + (tid >> 3) + 32 * i;
            const int x = m % W, y = (m / W) % H;
            uint32_t mk = 0;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const int yy = y + t9 / 3 - 1, xx = x + t9 % 3 - 1;
              if (m < M && yy >= 0 && yy < H && xx >= 0 && xx < W) mk |= 1u << t9;
            }
            vmask[i] = mk;
          }
          auto issue = [&](int s, int u) {   // warp 0, converged: slab of hidden rows + the slice's taps, one barrier
            __syncwarp();
            if (elect_one()) {
              const uint32_t dst = base + kOffSlab + (u & 1) * kSlab;
              mbar_arrive_expect_tx(slab_full(u & 1), slab_bytes + 1280u);
              tma_load_2d(dst, reinterpret_cast<const CUtensorMap*>(L->tm), (s0 + s) * 64, m0 - W - 1, slab_full(u & 1));
              bulk_g2s(dst + kSlabTaps, wdp + (size_t)(s0 + s) * 1280, 1280u, slab_full(u & 1));
            }
            __syncwarp();
          };
          if (warp == 0) issue(0, su);
          for (int s = 0; s < ns; ++s, ++su) {
            if (warp == 0 && s + 1 < ns) issue(s + 1, su + 1);   // the other buffer was released by the last bar_sync
            const int ch = (s0 + s) * 64 + c * 8;
            mbar_wait(slab_full(su & 1), (su >> 1) & 1);
            T(14);
            const uint8_t* slab = gbase + kOffSlab + (su & 1) * kSlab;
            const uint4* tp = reinterpret_cast<const uint4*>(slab + kSlabTaps + c * 160);
            // tap-major: one tap (LDS.128 broadcast) feeds the 4 rows of this thread -> 16 accumulators, few live registers
            const uint4 wb = tp[9];
            __nv_bfloat162 acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[i][q] = reinterpret_cast<const __nv_bfloat162*>(&wb)[q];
            const uint8_t* ctr0 = slab + ((tid >> 3) + W + 1) * 128 + c * 16;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const uint4 w = tp[t9];
              const __nv_bfloat162* pw = reinterpret_cast<const __nv_bfloat162*>(&w);
              const int toff = ((t9 / 3 - 1) * W + (t9 % 3 - 1)) * 128;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (vmask[i] >> t9 & 1) v = *reinterpret_cast<const uint4*>(ctr0 + i * 32 * 128 + toff);
                const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = __hfma2(pw[q], pv[q], acc[i][q]);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + (tid >> 3) + 32 * i;
              uint4 o;
              __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int q = 0; q < 4; ++q) po[q] = __hmax2(acc[i][q], __hmul2(acc[i][q], kslope));
              if (m < M) *reinterpret_cast<uint4*>(outp + (size_t)m * K + ch) = o;
            }
            bar_sync(2, kProducers);   // slab buffer free
            T(11);
          }
        } else {
          // stride 2: output (b, oy, ox) reads hidden pixels (2oy-1+ky, 2ox-1+kx); little reuse, taps straight from L2
          const int Wo = L->g.Wout, Ho = L->g.Hout;
          for (int s = 0; s < ns; ++s) {
            const int ch = (s0 + s) * 64 + c * 8;
            const uint4* tp = reinterpret_cast<const uint4*>(wdp + (size_t)(s0 + s) * 1280 + c * 160);
            uint4 wt[9];
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) wt[t9] = __ldg(tp + t9);
            const uint4 wb = __ldg(tp + 9);
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + (tid >> 3) + 32 * i;
              if (m >= M) continue;
              const int ox = m % Wo, t = m / Wo, oy = t % Ho, b = t / Ho;
              const __nv_bfloat16* ib = L->g.A + ((size_t)b * H * W) * K + ch;
              uint4 v[9];
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) {
                const int iy = 2 * oy - 1 + t9 / 3, ix = 2 * ox - 1 + t9 % 3;
                v[t9] = make_uint4(0, 0, 0, 0);
                if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                  v[t9] = __ldcg(reinterpret_cast<const uint4*>(ib + ((size_t)iy * W + ix) * K));
              }
              __nv_bfloat162 acc[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[q] = reinterpret_cast<const __nv_bfloat162*>(&wb)[q];
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) {
                const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v[t9]);
                const __nv_bfloat162* pw = reinterpret_cast<const __nv_bfloat162*>(&wt[t9]);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = __hfma2(pw[q], pv[q], acc[q]);
              }
              uint4 o;
              __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int q = 0; q < 4; ++q) po[q] = __hmax2(acc[q], __hmul2(acc[q], kslope));
              *reinterpret_cast<uint4*>(outp + (size_t)m * K + ch) = o;
            }
          }
        }
        // completion: the CTA-scope barrier orders every thread's stores before thread 0, whose gpu-scope release-add
        // on the row tile's counter is cumulative over them
        bar_sync(2, kProducers);
        if (tid == 0) red_release_add(cnt + S.cnt0 + mt, 1u);
        T(12);
        continue;
      }

      // ----------------- GEMM item with a gathered A operand: implicit-GEMM 3x3 taps, or upsample + concat ---------
      const int amode = S.amode;
      const int KB = S.KB;
      const GemmArgs& p = L->g;
      const int pM = p.M, pK = p.K, pCin = p.Cin, pHin = p.Hin, pWin = p.Win;
      int conv_pix[4], conv_yx[4];
      RowCoord rc{};
      if (amode == A_CONV3X3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int m = m0 + (tid >> 3) + 32 * i;
          int mm = m < pM ? m : 0;
          int ox = mm % p.Wout, t = mm / p.Wout;
          int oy = t % p.Hout, b = t / p.Hout;
          int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
          conv_pix[i] = (b * pHin + iy0) * pWin + ix0;
          conv_yx[i] = m < pM ? (((iy0 + 64) << 16) | (ix0 + 64)) : -1;
        }
      } else {
        int m = m0 + (tid >> 1);
        int mm = m < pM ? m : 0;
        int x = mm % p.Wout, t = mm / p.Wout;
        int y = t % p.Hout, b = t / p.Hout;
        // align_corners=True source coordinates (ATen area_pixel_compute_scale / upsample_bilinear2d)
        float sy = (float)(pHin - 1) / (float)(p.Hout - 1) * (float)y;
        float sx = (float)(pWin - 1) / (float)(p.Wout - 1) * (float)x;
        int y0 = (int)sy, x0 = (int)sx;
        int y1 = y0 + (y0 < pHin - 1 ? 1 : 0), x1 = x0 + (x0 < pWin - 1 ? 1 : 0);
        rc.wy1 = sy - (float)y0;
        rc.wy0 = 1.f - rc.wy1;
        rc.wx1 = sx - (float)x0;
        rc.wx0 = 1.f - rc.wx1;
        const __nv_bfloat16* fb = p.A + (size_t)b * pHin * pWin * pCin;
        rc.p00 = fb + (size_t)(y0 * pWin + x0) * pCin;
        rc.p01 = fb + (size_t)(y0 * pWin + x1) * pCin;
        rc.p10 = fb + (size_t)(y1 * pWin + x0) * pCin;
        rc.p11 = fb + (size_t)(y1 * pWin + x1) * pCin;
      }
      for (int kb = 0; kb < KB; ++kb) {
        const int jj = j + kb, s = jj % kS;
        mbar_wait(a_go(s), (gph >> s) & 1u);
        gph ^= 1u << s;
        const uint32_t a_s = base + s * kStage;
        if (amode == A_UPCAT) {
          const int r = tid >> 1, m = m0 + r;
          const int c2 = pK - pCin;
#pragma unroll
          for (int c = (tid & 1) * 4; c < (tid & 1) * 4 + 4; ++c) {
            const int k = (kb * 8 + c) * 8;
            const uint32_t dst = a_s + sw128_off(r, c);
            if (m < pM && k < pCin) {
              uint4 a = __ldcg(reinterpret_cast<const uint4*>(rc.p00 + k));
              uint4 b = __ldcg(reinterpret_cast<const uint4*>(rc.p01 + k));
              uint4 cc = __ldcg(reinterpret_cast<const uint4*>(rc.p10 + k));
              uint4 d = __ldcg(reinterpret_cast<const uint4*>(rc.p11 + k));
              uint4 o = lerp8(a, b, cc, d, rc);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w));
            } else {
              bool valid = m < pM && k < pK;
              const __nv_bfloat16* src = valid ? p.A2 + (size_t)m * c2 + (k - pCin) : p.A2;
              cp_async16(dst, src, valid);
            }
          }
        } else {  // implicit GEMM over the 9 taps of a dense 3x3 conv: k = tap*Cin + ci
          const int c = tid & 7;
          const int k = (kb * 8 + c) * 8;
          const int tap = k / pCin, ci = k - tap * pCin;
          const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 3) + 32 * i;
            const int iy = (conv_yx[i] >> 16) - 64 + ky, ix = (conv_yx[i] & 0xFFFF) - 64 + kx;
            bool valid = conv_yx[i] >= 0 && k < pK && iy >= 0 && iy < pHin && ix >= 0 && ix < pWin;
            const __nv_bfloat16* src = valid ? p.A + (size_t)(conv_pix[i] + ky * pWin + kx) * pCin + ci : p.A;
            cp_async16(a_s + sw128_off(r, c), src, valid);
          }
        }
        cp_async_commit();
        if (kb >= kLag) {   // publish the stage issued kLag k-blocks ago
          cp_async_wait<kLag>();
          fence_proxy_async();
          bar_sync(2, kProducers);
          if (tid == 0) mbar_arrive(full((jj - kLag) % kS));
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      bar_sync(2, kProducers);
      if (tid == 0)
        for (int q = (KB > kLag ? KB - kLag : 0); q < KB; ++q) mbar_arrive(full((j + q) % kS));
      j += KB;
      T(13);
    }
  } else if (warp == 8) {
    // ========================== loader (one elected lane): B tiles, and A tiles of plain operands ==================
    int j = 0, li = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      while (it >= tab[li].item_end) ++li;
      const LayerS& S = tab[li];
      if (S.kind != CK_GEMM) continue;
      const ChainLayer* __restrict__ L = layers + li;
      const int local = it - S.item0;
      const int NT = S.NT, BN = S.BN, KB = S.KB;
      const int mt = local / NT, nt = local - mt * NT;
      const int N = NT * BN;
      const uint8_t* Wp = L->g.W + (size_t)nt * BN * 128;
      const bool plain = S.amode == A_PLAIN;
      T(15);
      if (plain) {   // weights need no dependency, but the stage ring is shared: wait once per item, up front
        if (elect_one()) wait_deps(tab, S, mt * kBM, cnt);
        __syncwarp();
      }
      T(0);
      for (int kb = 0; kb < KB; ++kb, ++j) {
        const int s = j % kS;
        mbar_wait(empty(s), ((j / kS) & 1) ^ 1);
        T(1);
        if (elect_one()) {
          const uint32_t st_s = base + s * kStage;
          if (!plain) mbar_arrive(a_go(s));   // the workers may now fill the A half of this stage
          mbar_arrive_expect_tx(full(s), BN * 128);
          bulk_g2s(st_s + kABytes, Wp + (size_t)kb * N * 128, BN * 128, full(s));
          if (plain) {   // one 2-D TMA tensor copy (box 64 x 128, rows beyond M zero-filled)
            mbar_arrive_expect_tx(full(s), kABytes);
            tma_load_2d(st_s, reinterpret_cast<const CUtensorMap*>(L->tm), kb * 64, mt * kBM, full(s));
          }
        }
        __syncwarp();
        T(2);
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer (one elected lane of a converged warp) ======================
    int j = 0, t = 0, li = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      while (it >= tab[li].item_end) ++li;
      const LayerS& S = tab[li];
      if (S.kind != CK_GEMM) continue;
      const int KB = S.KB;
      const uint32_t idesc = umma_idesc_bf16(kBM, (uint32_t)S.BN);
      const int ab = t & 1;
      T(15);
      mbar_wait(acc_empty(ab), ((t >> 1) & 1) ^ 1);
      T(3);
      tc_fence_after();
      const uint32_t d = tmem + ab * kAccCols;
      for (int kb = 0; kb < KB; ++kb, ++j) {
        const int s = j % kS;
        mbar_wait(full(s), (j / kS) & 1);
        T(4);
        tc_fence_after();
        const uint32_t a_s = base + s * kStage;
        const uint64_t adesc = umma_desc_sw128(a_s), bdesc = umma_desc_sw128(a_s + kABytes);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kb | ks) != 0);
          umma_commit(empty(s));
          if (kb == KB - 1) umma_commit(acc_full(ab));
        }
        __syncwarp();
        T(5);
      }
      ++t;
    }
  } else {
    // ======================================= epilogue (warps 10-17) =================================================
    const int ew = warp - 10;
    const int lg = warp & 3;                 // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                  // column group of this warp
    float* const evec = reinterpret_cast<float*>(gbase + kOffEvec);
    const int et = ew * 32 + lane;
    uint8_t* const stg = gbase + kOffStg + ew * 2048;
    int t = 0, li = 0;
    for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
      while (it >= tab[li].item_end) ++li;
      const LayerS& S = tab[li];
      if (S.kind != CK_GEMM) continue;
      const ChainLayer* __restrict__ L = layers + li;
      const GemmArgs& p = L->g;
      const int local = it - S.item0;
      const int NT = S.NT, BN = S.BN;
      const int mt = local / NT, nt = local - mt * NT;
      const int n0 = nt * BN, m0 = mt * kBM;
      const int ncg = BN >= 64 ? 2 : 1;
      const int cols_per = BN / ncg;
      const int ab = t & 1;
      const int pM = p.M, ldc = p.ldc, leaky_on = p.leaky;
      const __nv_bfloat16* const res_pre = p.res_pre;
      const __nv_bfloat16* const res_post = p.res_post;
      const float* const post_scale = p.post_scale;
      __nv_bfloat16* const Cp = p.C;
      __nv_bfloat16* const vt = p.vt;
      const int vt_col0 = p.vt_col0;
      float* const ev = evec + (t & 1) * 1024;
      if (et < BN) {
        ev[et] = __ldg(p.bias + n0 + et);
        if (res_pre) ev[256 + et] = __ldg(p.rscale + n0 + et);
        if (post_scale) {
          ev[512 + et] = __ldg(post_scale + n0 + et);
          ev[768 + et] = __ldg(p.post_shift + n0 + et);
        }
      }
      T(15);
      bar_sync(1, 256);
      T(6);
      const int m = m0 + lg * 32 + lane;
      const bool row_ok = m < pM;
      const bool active = cg < ncg;
      const int cbeg = cg * cols_per, cend = active ? cbeg + cols_per : cbeg;
      const __nv_bfloat16* rsrc = res_pre ? res_pre + (size_t)m * p.ld_rpre : res_post ? res_post + (size_t)m * p.ld_rpost : nullptr;
      uint4 rnext[4];
      auto fetch_res = [&](int c0) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          rnext[g] = (rsrc && row_ok) ? __ldcg(reinterpret_cast<const uint4*>(rsrc + n0 + c0 + 8 * g)) : make_uint4(0, 0, 0, 0);
      };
      // residual rows: requested before the accumulator wait (latency hidden behind the main loop) unless they are
      // produced inside this program -- then only the finished accumulator proves that they are complete
      const bool res_late = L->res_late != 0;
      if (cbeg < cend && !res_late) fetch_res(cbeg);
      mbar_wait(acc_full(ab), (t >> 1) & 1);
      T(7);
      tc_fence_after();
      if (cbeg < cend && res_late) fetch_res(cbeg);
      const uint32_t trow = tmem + ab * kAccCols + ((uint32_t)(lg * 32) << 16);
      const bool transposed = vt && n0 >= vt_col0;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(trow + c0, acc);
        uint4 rcur[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
        if (c0 + 32 < cend) fetch_res(c0 + 32);
        tmem_ld_wait32(acc);
        if (row_ok) {
          const int n = n0 + c0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 8 columns per group -> one 16 B store
            float v[8];
            const float4 b0 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g);
            const float4 b1 = *reinterpret_cast<const float4*>(ev + c0 + 8 * g + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(acc[8 * g + q]) + bb[q];
            if (res_pre) {
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
            if (leaky_on) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], kLeaky * v[q]);
            }
            if (res_post) {
              const uint32_t* pr = &rcur[g].x;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] += bf16_lo(pr[q]);
                v[2 * q + 1] += bf16_hi(pr[q]);
              }
            }
            if (post_scale) {
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
            if (transposed) {   // transposed store of a value projection (see GemmArgs::vt)
              const int frame = m / 100, key = m - frame * 100;
              const int colv = n + 8 * g - vt_col0;   // j * 512 + channel
              __nv_bfloat16* dst = vt + ((size_t)frame * 2048 + colv) * 128 + key;
#pragma unroll
              for (int q = 0; q < 8; ++q) dst[q * 128] = __float2bfloat16_rn(v[q]);
            } else {
              uint4 o;
              o.x = pack_bf16(v[0], v[1]);
              o.y = pack_bf16(v[2], v[3]);
              o.z = pack_bf16(v[4], v[5]);
              o.w = pack_bf16(v[6], v[7]);
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = o;
            }
          }
        }
        if (!transposed) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + (lane >> 2), c = lane & 3;
            const uint4 o = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((c ^ ((rl >> 1) & 3)) << 4));
            const int ml = m0 + lg * 32 + rl;
            if (ml < pM) *reinterpret_cast<uint4*>(Cp + (size_t)ml * ldc + n0 + c0 + 8 * c) = o;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));
      T(8);
      // completion of (layer, row tile, this column tile): the barrier orders every epilogue thread's stores before
      // thread 0, whose gpu-scope release-add is cumulative over them; the row tile is complete when all NT column
      // tiles have arrived
      bar_sync(1, 256);
      if (et == 0) red_release_add(cnt + S.cnt0 + mt, 1u);
      T(9);
      ++t;
    }
  }

  T(15);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

unsigned long long* g_chain_dbg = nullptr;

}  // namespace

void chain_dbg_report() {
  if (!g_chain_dbg) return;
  unsigned long long h[16] = {0};
  cudaDeviceSynchronize();
  cudaMemcpy(h, g_chain_dbg, 128, cudaMemcpyDeviceToHost);
  const char* names[16] = {"L:wait_deps", "L:wait_empty", "L:issue", "M:wait_acc_empty", "M:wait_full", "M:issue",
                           "E:stage+bar", "E:wait_acc_full", "E:drain+store", "E:signal", "W:wait_deps", "W:dw_compute",
                           "W:dw_done", "W:gather", "W:wait_slab", "other"};
  fprintf(stderr, "[casync chain dbg] cycles summed over CTAs and launches:\n   ");
  for (int i = 0; i < 16; ++i) fprintf(stderr, " %s=%.3g", names[i], (double)h[i]);
  fprintf(stderr, "\n");
}

int chain_init() {
  if (const char* c = getenv("CASYNC_CHAIN_DBG"))
    if (atoi(c) > 0 && !g_chain_dbg && cudaMalloc(&g_chain_dbg, 128) == cudaSuccess) cudaMemset(g_chain_dbg, 0, 128);
  return (int)cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
}

// ---------------------------------------------------------------------------------------------------------------------
static int chain_max_layers() {   // developer switch: CASYNC_CHAIN_MAXL caps the layers per program
  static const int v = [] {
    const char* e = getenv("CASYNC_CHAIN_MAXL");
    const int n = e ? atoi(e) : 0;
    return n > 0 && n < kChainMaxLayers ? n : kChainMaxLayers;
  }();
  return v;
}

void Chain::reset() {
  layers_.clear();
  outs_.clear();
  items_ = 0;
  counters_ = 0;
}

// Which earlier layers of the program wrote what this operand reads?  Operand = rows of `ncols` elements at `ptr` with
// pitch `ld`, `bytes` in total.  Returns -1 when more than two producers are involved.
int Chain::find_deps(ChainLayer& L, const void* ptr, int ld, int ncols, size_t bytes, int mode, int a, int b) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(ptr);
  for (const Out& o : outs_) {
    if (p + bytes <= o.base || o.base + o.bytes <= p) continue;   // disjoint byte ranges
    if (o.ld == ld && o.ld > 0) {
      // same pitch: the column intervals may still be disjoint (the two halves of the `cat` buffer).  Column of the
      // operand's first element relative to the producer's row origin:
      const long d = (long)(p - o.base) / 2;
      const long col = ((d % ld) + ld) % ld;
      if (col + ncols <= ld && col >= o.ncols) continue;
    }
    bool dup = false;
    for (int d = 0; d < 2; ++d)
      if (L.dep[d].mode && L.dep[d].layer == o.layer) dup = true;
    if (dup) continue;
    int slot = L.dep[0].mode == 0 ? 0 : L.dep[1].mode == 0 ? 1 : -1;
    if (slot < 0) return -1;
    L.dep[slot].mode = mode;
    L.dep[slot].layer = o.layer;
    L.dep[slot].a = a;
    L.dep[slot].b = b;
    // a "same rows" dependency needs the same row space
    if (mode == 1 && layers_[o.layer].out_rows != L.out_rows) return -1;
  }
  return 0;
}

int Chain::add_gemm(const GemmArgs& g) {
  if ((int)layers_.size() >= chain_max_layers()) return -1;
  if (g.M <= 0 || g.N % 32 != 0 || g.K % 64 != 0) return -1;
  ChainLayer L;
  memset(&L, 0, sizeof L);
  L.g = g;
  L.g.dbg = nullptr;
  L.kind = CK_GEMM;
  int bn = 32;
  for (int c : {64, 128, 192, 256})
    if (g.N % c == 0 && !(g.vt && g.vt_col0 % c)) bn = c;
  L.BN = bn;
  L.NT = g.N / bn;
  L.MT = (g.M + kBM - 1) / kBM;
  L.need = L.NT;
  L.out_rows = g.M;
  int e = 0;
  if (g.amode == A_PLAIN) {
    if (gemm_encode_map(L.tm, g.A, g.M, g.K, g.lda, kBM, true)) return -1;
    e = find_deps(L, g.A, g.lda, g.K, ((size_t)(g.M - 1) * g.lda + g.K) * 2, 1, 0, 0);
  } else if (g.amode == A_CONV3X3) {
    const int frames = g.M / (g.Hout * g.Wout);
    e = find_deps(L, g.A, g.Cin, g.Cin, (size_t)frames * g.Hin * g.Win * g.Cin * 2, 2, g.Hout * g.Wout, g.Hin * g.Win);
  } else {
    const int frames = g.M / (g.Hout * g.Wout), c2 = g.K - g.Cin;
    e = find_deps(L, g.A, g.Cin, g.Cin, (size_t)frames * g.Hin * g.Win * g.Cin * 2, 2, g.Hout * g.Wout, g.Hin * g.Win);
    if (!e) e = find_deps(L, g.A2, c2, c2, (size_t)g.M * c2 * 2, 1, 0, 0);
  }
  if (e) return -1;
  {
    const __nv_bfloat16* r = g.res_pre ? g.res_pre : g.res_post;
    const int ldr = g.res_pre ? g.ld_rpre : g.ld_rpost;
    if (r) {
      const unsigned char* rp = reinterpret_cast<const unsigned char*>(r);
      const size_t rbytes = ((size_t)(g.M - 1) * ldr + g.N) * 2;
      for (const Out& o : outs_)
        if (!(rp + rbytes <= o.base || o.base + o.bytes <= rp)) L.res_late = 1;
    }
  }
  L.item0 = items_;
  L.items = L.MT * L.NT;
  L.cnt0 = counters_;
  const int idx = (int)layers_.size();
  layers_.push_back(L);
  items_ += L.items;
  counters_ += L.MT;
  outs_.push_back({reinterpret_cast<const unsigned char*>(g.C), ((size_t)(g.M - 1) * g.ldc + g.N) * 2, g.ldc, g.N, idx});
  return 0;
}

int Chain::add_dw(const __nv_bfloat16* in, __nv_bfloat16* out, const uint8_t* wdp, int batch, int H, int W, int C,
                  int stride) {
  if ((int)layers_.size() >= chain_max_layers()) return -1;
  if (C % 64 != 0 || H != W || (stride != 1 && stride != 2) || !wdp) return -1;
  if (stride == 1 && 130 + 2 * W > kSlabRows) return -1;
  const int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? W / 2 : W;
  ChainLayer L;
  memset(&L, 0, sizeof L);
  L.kind = CK_DW;
  L.g.A = in;
  L.g.C = out;
  L.g.W = wdp;
  L.g.K = C;
  L.g.N = C;
  L.g.M = batch * Ho * Wo;
  L.g.Hin = H;
  L.g.Win = W;
  L.g.Hout = Ho;
  L.g.Wout = Wo;
  L.g.stride = stride;
  L.g.ldc = C;
  L.NT = (C / 64 + 3) / 4;
  L.MT = (L.g.M + kBM - 1) / kBM;
  L.need = L.NT;
  L.out_rows = L.g.M;
  const int rows_in = batch * H * W;
  if (stride == 1 && gemm_encode_map(L.tm, in, rows_in, C, C, 130 + 2 * W, false)) return -1;
  if (find_deps(L, in, C, C, (size_t)rows_in * C * 2, stride == 1 ? 1 : 2, stride == 1 ? W + 1 : Ho * Wo, H * W)) return -1;
  L.item0 = items_;
  L.items = L.MT * L.NT;
  L.cnt0 = counters_;
  const int idx = (int)layers_.size();
  layers_.push_back(L);
  items_ += L.items;
  counters_ += L.MT;
  outs_.push_back({reinterpret_cast<const unsigned char*>(out), (size_t)L.g.M * C * 2, C, C, idx});
  return 0;
}

size_t Chain::scratch_bytes() const {
  return layers_.size() * sizeof(ChainLayer) + (((size_t)counters_ * 4 + 255) & ~(size_t)255);
}

int Chain::launch(void* scratch, size_t scratch_cap, std::vector<unsigned char>& last, cudaStream_t st) {
  if (layers_.empty()) return 0;
  const size_t lbytes = layers_.size() * sizeof(ChainLayer);
  const size_t total = scratch_bytes();
  if (!scratch || ((uintptr_t)scratch & 255) || total > scratch_cap) {
    reset();
    return (int)cudaErrorInvalidValue;
  }
  cudaError_t e = cudaSuccess;
  // the descriptors only change with the batch size / workspace address: upload when they differ from the device copy
  if (last.size() != lbytes || memcmp(last.data(), layers_.data(), lbytes) != 0) {
    last.assign(reinterpret_cast<const unsigned char*>(layers_.data()),
                reinterpret_cast<const unsigned char*>(layers_.data()) + lbytes);
    // pageable source: staged by the runtime before the call returns; ordered after earlier work of the stream
    e = cudaMemcpyAsync(scratch, last.data(), lbytes, cudaMemcpyHostToDevice, st);
  }
  unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(scratch) + lbytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(counters, 0, (size_t)counters_ * 4, st);
  if (e == cudaSuccess) {
    const int sms = gemm_num_sms();
    const int grid = items_ < sms ? items_ : sms;
    e = launch_pdl(chain_kernel, dim3(grid), dim3(kThreads), kSmem, st, reinterpret_cast<const ChainLayer*>(scratch),
                   (int)layers_.size(), items_, counters, g_chain_dbg);
  }
  reset();
  return (int)e;
}

}  // namespace casync
