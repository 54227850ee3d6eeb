// Fused InvertedResidual (module/unet.py:8-40) for the high-resolution, HBM-bound blocks:
//   pw1 (tcgen05) -> BN+leaky -> depthwise 3x3 (CUDA cores, from shared memory) -> BN+leaky -> pw2 (tcgen05)
//   -> BN+leaky (+ skip)            -- the x2-expanded hidden tensor never leaves the SM.
//
// One persistent CTA per SM walks over 2-D patches of one frame.  A patch is a 16x16 window of HIDDEN pixels
// (= two 128-row UMMA tiles): stride 1 -> 14x14 outputs, stride 2 -> 7x7 outputs (15x15 of the window used).
// Per patch:
//   A1   [256 px x Cin]  bf16, SWIZZLE_128B K-major, produced by cp.async (plain NHWC rows) or computed on the
//        fly (decoder: bilinear x2 of the low-res tensor for the first Cin/2 channels, skip tensor for the rest,
//        module/unet.py:90-96);
//   D1   = A1 . W1^T      two M=128 tcgen05.mma groups into TMEM (fp32);
//   HID  = leaky(D1 + b1) bf16 in shared memory (same swizzled layout), forced to 0 outside the image because the
//        depthwise conv zero-pads the hidden tensor (module/unet.py:21-27);
//   A2   = leaky(dw3x3(HID) + bd)  bf16, again the UMMA A layout (rows = output pixels of the patch);
//   D2   = A2 . W2^T      tcgen05.mma into TMEM;   out = leaky(D2 + b2) (+ x)  -> global NHWC bf16.
// Weights (W1, W2 tiles in their packed swizzled image, depthwise taps, biases) are fetched once per CTA with
// bulk-async copies and stay resident in shared memory.
#include "fused_ir.cuh"

#include <cuda_bf16.h>

namespace casync {

namespace {

constexpr int kThreads = 512;
constexpr int kTileBytes = 128 * 128;       // one 128-row x 128 B swizzled tile
constexpr int kKbBytes = 2 * kTileBytes;    // 256 rows (both M tiles) of one 64-channel k-block

template <int CIN, int HC, int COUT, int STRIDE>
struct FCfg {
  static constexpr int KB1 = (CIN + 63) / 64;   // A1 k-blocks
  static constexpr int KB2 = HC / 64;           // hidden k-blocks
  static constexpr int T = STRIDE == 1 ? 14 : 7;
  static constexpr int A2_TILES = STRIDE == 1 ? 2 : 1;
  static constexpr int oA1 = 0;
  static constexpr int oHID = oA1 + KB1 * kKbBytes;
  static constexpr int oA2 = oHID + KB2 * kKbBytes;
  static constexpr int oW1 = oA2 + KB2 * kKbBytes;
  static constexpr int oW2 = oW1 + KB1 * HC * 128;
  static constexpr int oWD = oW2 + KB2 * COUT * 128;
  static constexpr int oB1 = oWD + 9 * HC * 4;
  static constexpr int oBD = oB1 + HC * 4;
  static constexpr int oB2 = oBD + HC * 4;
  static constexpr int oBAR = oB2 + COUT * 4;
  static constexpr int kSmem = oBAR + 64 + 1024;
  static constexpr uint32_t kWeightBytes = KB1 * HC * 128 + KB2 * COUT * 128 + 9 * HC * 4 + 2 * HC * 4 + COUT * 4;
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int CIN, int HC, int COUT, int STRIDE, bool UPCAT, bool RES>
__global__ void __launch_bounds__(kThreads, 1) fused_ir_kernel(const FusedArgs p) {
  using C = FCfg<CIN, HC, COUT, STRIDE>;
  constexpr int KB1 = C::KB1, KB2 = C::KB2, T = C::T;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA1 = base + C::oA1, sHID = base + C::oHID, sA2 = base + C::oA2, sW1 = base + C::oW1,
                 sW2 = base + C::oW2, sWD = base + C::oWD, sB1 = base + C::oB1, sBD = base + C::oBD,
                 sB2 = base + C::oB2, sBAR = base + C::oBAR;
  const uint32_t barW = sBAR, bar1 = sBAR + 8, bar2 = sBAR + 16, tmem_slot = sBAR + 24;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(barW, 1);
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(barW, C::kWeightBytes);
    for (int kb = 0; kb < KB1; ++kb)   // chunk rows [0,HC) of every k-block of W1
      bulk_g2s(sW1 + kb * HC * 128, p.W1 + (size_t)kb * (2 * CIN) * 128, HC * 128, barW);
    bulk_g2s(sW2, p.W2, KB2 * COUT * 128, barW);
    bulk_g2s(sWD, p.wd, 9 * HC * 4, barW);
    bulk_g2s(sB1, p.b1, HC * 4, barW);
    bulk_g2s(sBD, p.bd, HC * 4, barW);
    bulk_g2s(sB2, p.b2, COUT * 4, barW);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tD1 = tmem, tD2 = tmem + 2 * HC;
  mbar_wait(barW, 0);

  const int W = p.W, Wo = W / STRIDE;
  const int PD = (Wo + T - 1) / T;
  const int NP = p.batch * PD * PD;
  constexpr uint32_t idesc1 = umma_idesc_bf16(128, HC), idesc2 = umma_idesc_bf16(128, COUT);

  struct Patch {
    int b, OY0, OX0, GY0, GX0;
  };
  auto decode = [&](int patch) {
    Patch q;
    q.b = patch / (PD * PD);
    const int pr = patch - q.b * PD * PD, py = pr / PD, px = pr - py * PD;
    q.OY0 = py * T;
    q.OX0 = px * T;
    q.GY0 = q.OY0 * STRIDE - 1;   // global coords of hidden-window pixel (0,0)
    q.GX0 = q.OX0 * STRIDE - 1;
    return q;
  };

  // ---------------- A1: 256 window pixels x CIN channels; two threads per pixel row ---------------------------
  auto produce_a1 = [&](const Patch& q) {
    const int row = tid >> 1, half = tid & 1;
    const int gy = q.GY0 + (row >> 4), gx = q.GX0 + (row & 15);
    const bool inside = gy >= 0 && gy < W && gx >= 0 && gx < W;
    constexpr int NCH = CIN / 8;   // 16-byte chunks per row
    if constexpr (!UPCAT) {
      const __nv_bfloat16* src = p.in + ((size_t)(q.b * W + (inside ? gy : 0)) * W + (inside ? gx : 0)) * CIN;
#pragma unroll
      for (int c = half; c < NCH; c += 2)
        cp_async16(sA1 + (c >> 3) * kKbBytes + sw128_off(row, c & 7), src + c * 8, inside);
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
#pragma unroll
      for (int c = half; c < NCH; c += 2) {
        const uint32_t dst = sA1 + (c >> 3) * kKbBytes + sw128_off(row, c & 7);
        if (c * 8 < C1) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (inside) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p00 + c * 8));
            const uint4 bq = __ldg(reinterpret_cast<const uint4*>(p01 + c * 8));
            const uint4 cq = __ldg(reinterpret_cast<const uint4*>(p10 + c * 8));
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(p11 + c * 8));
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&bq);
            const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&cq);
            const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&d);
            __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              po[i] = __hfma2(w11, pd[i], __hfma2(w10, pc[i], __hfma2(w01, pb[i], __hmul2(w00, pa[i]))));
          }
          sts128(dst, o.x, o.y, o.z, o.w);
        } else {
          cp_async16(dst, sk + (c * 8 - C1), inside);
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
  };

  // ---------------- GEMM1: D1[256 x HC] = A1 . W1^T (thread 0) ---------------------------------------------------
  auto issue_gemm1 = [&]() {
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int kb = 0; kb < KB1; ++kb) {
        const uint64_t ad = umma_desc_sw128(sA1 + kb * kKbBytes + t * kTileBytes);
        const uint64_t bd = umma_desc_sw128(sW1 + kb * HC * 128);
        constexpr int KS_ALL = (CIN + 15) / 16;
        const int ks_n = (KS_ALL - kb * 4) < 4 ? (KS_ALL - kb * 4) : 4;
        for (int ks = 0; ks < ks_n; ++ks) umma_bf16(tD1 + t * HC, ad + 2 * ks, bd + 2 * ks, idesc1, (kb | ks) != 0);
      }
    }
    umma_commit(bar1);
  };

  const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);
  // ---------------- drain D1 -> HID (bias, leaky; 0 outside the image) ------------------------------------------
  auto drain1 = [&](const Patch& q, uint32_t parity) {
    mbar_wait(bar1, parity);
    tc_fence_after();
    const int lg = warp & 3, sub = warp >> 2;
    const int tile = sub >> 1, col0 = (sub & 1) * (HC / 2);
    const int row = tile * 128 + lg * 32 + lane;
    const int gy = q.GY0 + (row >> 4), gx = q.GX0 + (row & 15);
    const bool inside = gy >= 0 && gy < W && gx >= 0 && gx < W;
#pragma unroll
    for (int cc = 0; cc < HC / 2; cc += 32) {
      uint32_t acc[32];
      tmem_ld32(tD1 + tile * HC + col0 + cc + ((uint32_t)(lg * 32) << 16), acc);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int col = col0 + cc + g8 * 8;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
        if (inside) {
          const float4 ba = lds_f4(sB1 + col * 4), bb = lds_f4(sB1 + col * 4 + 16);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // bias in fp32, one rounding to bf16, LeakyReLU on the packed pair
            __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[g8 * 8 + 2 * j]) + bv[2 * j],
                                                     __uint_as_float(acc[g8 * 8 + 2 * j + 1]) + bv[2 * j + 1]);
            v = __hmax2(v, __hmul2(v, kslope));
            o[j] = *reinterpret_cast<uint32_t*>(&v);
          }
        }
        sts128(sHID + (col >> 6) * kKbBytes + sw128_off(row, (col & 63) >> 3), o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before();
  };

  // ---------------- depthwise 3x3 (packed bf16x2 FMAs) : HID -> A2 ------------------------------------------------
  // Thread = one 4-channel group (8 B).  The taps / bias of the group live in registers for the whole kernel.
  constexpr int G = HC / 4, IPW = 32 / G;
  const int dw_g = lane % G, dw_sub = lane / G, dw_ch = 4 * dw_g;
  __nv_bfloat162 wt[9][2], wbias[2];
#pragma unroll
  for (int t9 = 0; t9 < 9; ++t9) {
    const float4 w4 = lds_f4(sWD + (t9 * HC + dw_ch) * 4);
    wt[t9][0] = __floats2bfloat162_rn(w4.x, w4.y);
    wt[t9][1] = __floats2bfloat162_rn(w4.z, w4.w);
  }
  {
    const float4 b4 = lds_f4(sBD + dw_ch * 4);
    wbias[0] = __floats2bfloat162_rn(b4.x, b4.y);
    wbias[1] = __floats2bfloat162_rn(b4.z, b4.w);
  }
  auto dwconv = [&]() {
    const uint32_t hid_g = sHID + (dw_ch >> 6) * kKbBytes + (dw_g & 1) * 8;
    const uint32_t a2_g = sA2 + (dw_ch >> 6) * kKbBytes + (dw_g & 1) * 8;
    const uint32_t chunk = (dw_ch & 63) >> 3;
    auto hid_at = [&](int hy, int hx) { return lds64(hid_g + sw128_off(hy * 16 + hx, chunk)); };
    auto tap = [&](__nv_bfloat162* a, const uint2& v, const __nv_bfloat162* w) {
      a[0] = __hfma2(w[0], *reinterpret_cast<const __nv_bfloat162*>(&v.x), a[0]);
      a[1] = __hfma2(w[1], *reinterpret_cast<const __nv_bfloat162*>(&v.y), a[1]);
    };
    auto finish = [&](__nv_bfloat162* a, uint32_t addr) {
      a[0] = __hmax2(a[0], __hmul2(a[0], kslope));
      a[1] = __hmax2(a[1], __hmul2(a[1], kslope));
      sts64(addr, *reinterpret_cast<uint32_t*>(&a[0]), *reinterpret_cast<uint32_t*>(&a[1]));
    };
    if constexpr (STRIDE == 1) {
      // item = (output column ox, half of the 14 output rows); vertical sliding 3x3 window.  Window pixel
      // (hy,hx) is row hy*16+hx of the swizzled tile: its XOR term depends on hx only, so the three column
      // bases are computed once per item and rows are reached with immediate offsets (2048 B per window row).
      for (int item = warp * IPW + dw_sub; item < 28; item += 16 * IPW) {
        const int ox = item >> 1, oy0 = (item & 1) * 7;
        uint32_t cb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) cb[k] = hid_g + oy0 * 2048 + (ox + k) * 128 + ((chunk ^ ((ox + k) & 7)) << 4);
        uint2 r0[3], r1[3], r2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          r0[k] = lds64(cb[k]);
          r1[k] = lds64(cb[k] + 2048);
        }
        int op = oy0 * 14 + ox;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
#pragma unroll
          for (int k = 0; k < 3; ++k) r2[k] = lds64(cb[k] + (i + 2) * 2048);
          __nv_bfloat162 a[2] = {wbias[0], wbias[1]};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            tap(a, r0[k], wt[k]);
            tap(a, r1[k], wt[3 + k]);
            tap(a, r2[k], wt[6 + k]);
          }
          finish(a, a2_g + op * 128 + ((chunk ^ (op & 7)) << 4));   // tiles are contiguous: row op of A2
          op += 14;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            r0[k] = r1[k];
            r1[k] = r2[k];
          }
        }
      }
    } else {
      // stride 2: 7x7 outputs, centre of output (oy,ox) is window pixel (2oy+1, 2ox+1)
      for (int item = warp * IPW + dw_sub; item < 49; item += 16 * IPW) {
        const int oy = item / 7, ox = item - oy * 7;
        __nv_bfloat162 a[2] = {wbias[0], wbias[1]};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) tap(a, hid_at(2 * oy + ky, 2 * ox + kx), wt[ky * 3 + kx]);
        finish(a, a2_g + sw128_off(item, chunk));
      }
    }
    fence_proxy_async();
  };

  // ---------------- GEMM2: D2[out px x COUT] = A2 . W2^T (thread 0) ----------------------------------------------
  auto issue_gemm2 = [&]() {
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < C::A2_TILES; ++t) {
#pragma unroll
      for (int kb = 0; kb < KB2; ++kb) {
        const uint64_t ad = umma_desc_sw128(sA2 + kb * kKbBytes + t * kTileBytes);
        const uint64_t bd = umma_desc_sw128(sW2 + kb * COUT * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tD2 + t * COUT, ad + 2 * ks, bd + 2 * ks, idesc2, (kb | ks) != 0);
      }
    }
    umma_commit(bar2);
  };

  // ---------------- epilogue: D2 -> leaky(+b2) (+skip) -> global NHWC bf16 ----------------------------------------
  auto epilogue = [&](const Patch& q, uint32_t parity) {
    mbar_wait(bar2, parity);
    tc_fence_after();
    const int lg = warp & 3, sub = warp >> 2;
    const int tile = sub >> 1, col0 = (sub & 1) * (COUT / 2);
    if (tile < C::A2_TILES) {
      const int op = tile * 128 + lg * 32 + lane;
      const int oy = op / T, ox = op - oy * T;
      const int gy = q.OY0 + oy, gx = q.OX0 + ox;
      const bool valid = op < T * T && gy < Wo && gx < Wo;
      const size_t opix = (size_t)(q.b * Wo + (valid ? gy : 0)) * Wo + (valid ? gx : 0);
#pragma unroll
      for (int cc = 0; cc < COUT / 2; cc += 16) {
        uint32_t acc[16];
        tmem_ld16(tD2 + tile * COUT + col0 + cc + ((uint32_t)(lg * 32) << 16), acc);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const int col = col0 + cc + g8 * 8;
            const float4 ba = lds_f4(sB2 + col * 4), bb = lds_f4(sB2 + col * 4 + 16);
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = leaky(__uint_as_float(acc[g8 * 8 + j]) + bv[j]);
            if constexpr (RES) {   // stride 1, CIN == COUT: skip = the block input at the same pixel
              const uint4 r = __ldg(reinterpret_cast<const uint4*>(p.in + opix * CIN + col));
              const uint32_t* pr = &r.x;
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
    }
    tc_fence_before();
  };

  // ---------------- software pipeline over this CTA's patches ------------------------------------------------------
  // steady state per patch i:  dw(i) | GEMM2(i) issued | drain(i+1) + epilogue(i) + A1(i+2) | GEMM1(i+2) issued
  // so both MMA groups run underneath CUDA-core phases of neighbouring patches.
  long long tphase = 0;
  auto mark = [&](int ph) {   // developer phase timing: thread 0 of every CTA accumulates cycles per phase
    if (p.dbg && tid == 0) {
      const long long now = clock64();
      if (ph >= 0) atomicAdd(p.dbg + ph, (unsigned long long)(now - tphase));
      tphase = now;
    }
  };
  const int first = blockIdx.x, step = gridDim.x;
  const int n_mine = first < NP ? (NP - first + step - 1) / step : 0;
  if (n_mine > 0) {
    Patch cur = decode(first), nxt = cur;
    produce_a1(cur);
    __syncthreads();
    if (tid == 0) issue_gemm1();
    drain1(cur, 0);
    if (n_mine > 1) {
      nxt = decode(first + step);
      produce_a1(nxt);      // GEMM1(0) finished (bar1 waited in drain1): A1 is free
    }
    __syncthreads();
    if (n_mine > 1 && tid == 0) issue_gemm1();
    mark(-1);
    for (int i = 0; i < n_mine; ++i) {
      dwconv();             // HID(i) -> A2
      mark(0);
      __syncthreads();
      mark(1);
      if (tid == 0) issue_gemm2();
      if (i + 1 < n_mine) drain1(nxt, (i + 1) & 1);      // HID(i+1) <- D1(i+1)
      mark(2);
      epilogue(cur, i & 1);                              // D2(i) -> global
      mark(3);
      cur = nxt;
      if (i + 2 < n_mine) {
        nxt = decode(first + (i + 2) * step);
        produce_a1(nxt);
      }
      mark(4);
      __syncthreads();
      mark(5);
      if (i + 2 < n_mine && tid == 0) issue_gemm1();
      mark(6);
    }
  }

  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int CIN, int HC, int COUT, int STRIDE, bool UPCAT, bool RES>
int launch_t(const FusedArgs& a, cudaStream_t st) {
  using C = FCfg<CIN, HC, COUT, STRIDE>;
  static bool attr_set = false;
  auto kfn = fused_ir_kernel<CIN, HC, COUT, STRIDE, UPCAT, RES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int Wo = a.W / STRIDE, PD = (Wo + C::T - 1) / C::T;
  const int NP = a.batch * PD * PD;
  const int grid = NP < a.num_sms ? NP : a.num_sms;
  kfn<<<grid, kThreads, C::kSmem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

bool fused_ir_supported(int cin, int cout, int stride, bool upcat, bool res) {
  FusedArgs a{};
  a.cin = cin; a.cout = cout; a.stride = stride; a.upcat = upcat; a.res = res; a.batch = 0;
  return launch_fused_ir(a, nullptr) != -1;
}

// returns -1 when the block shape has no fused instantiation, else 0 / cudaError
int launch_fused_ir(const FusedArgs& a, cudaStream_t st) {
#define CASE(CIN_, COUT_, S_, U_, R_)                                                         \
  if (a.cin == CIN_ && a.cout == COUT_ && a.stride == S_ && a.upcat == U_ && a.res == R_) {   \
    if (a.batch <= 0) return 0;                                                               \
    return launch_t<CIN_, 2 * CIN_, COUT_, S_, U_, R_>(a, st);                                \
  }
  CASE(32, 64, 2, false, false)    // down1.0
  CASE(64, 64, 1, false, true)     // down1.1, up2.1
  CASE(64, 128, 2, false, false)   // down2.0
  CASE(64, 32, 1, true, false)     // up4.0
  CASE(32, 32, 1, false, true)     // up4.1, up3.1
  CASE(32, 64, 1, false, false)    // audio conv1
  CASE(64, 128, 1, false, false)   // audio conv2
#undef CASE
  return -1;
}

}  // namespace casync
