// tcgen05 GEMM kernel (sm_100a).  One CTA = one 128 x BN output tile:
//   * A tile (128 rows x 64 bf16 per k-block) is produced into SWIZZLE_128B shared memory by all 128
//     threads: cp.async 16 B chunks (plain rows / implicit-GEMM 3x3 taps) or computed on the fly
//     (bilinear x2 upsample of the low-res tensor for the decoder's concat, module/unet.py:90-96);
//   * the weight tile (BN rows) is already stored in global memory as the swizzled shared-memory image,
//     so one thread fetches it with a single bulk-async (TMA engine) copy completing on an mbarrier;
//   * thread 0 issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block into a TMEM accumulator and
//     commits to an mbarrier that releases the stage;
//   * 4 warps drain TMEM (tcgen05.ld 32x32b) and apply the fused epilogue (folded-BN bias, LeakyReLU,
//     pre/post residuals, trailing BN) before 16-byte bf16 stores.
// Several CTAs are resident per SM (TMEM columns = BN <= 256), which overlaps one CTA's epilogue with
// another's main loop.
#include "gemm_tc.cuh"

namespace casync {

namespace {

constexpr int kBM = 128;
constexpr int kABytes = kBM * 128;  // 16 KiB: 128 rows x 128 B

template <int BN, int S>
struct Cfg {
  static constexpr int kStage = kABytes + BN * 128;
  static constexpr int kSmem = S * kStage + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

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

template <int BN, int S>
__global__ void __launch_bounds__(128) gemm_tc_kernel(const GemmArgs p) {
  using C = Cfg<BN, S>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + S * C::kStage;
  const uint32_t tmem_slot = bar_base + 16 * S;
  auto full_b = [&](int s) { return bar_base + 8u * s; };
  auto mma_done = [&](int s) { return bar_base + 8u * (S + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int KB = (p.K + 63) >> 6;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_b(s), 1);
      mbar_init(mma_done(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  // ---- per-thread A-row bookkeeping -------------------------------------------------------------
  // PLAIN / CONV3X3: thread owns chunk (tid & 7) of rows (tid >> 3) + 16 i, i = 0..7 (a warp copies four
  // full 128 B rows per instruction).  UPCAT: thread owns row tid (needs 4 taps + weights per row).
  int conv_pix[8];   // CONV3X3: (b*Hin + iy0)*Win + ix0 for tap (0,0)
  int conv_yx[8];    // CONV3X3: iy0 (hi 16) | ix0 (lo 16), biased by +64
  RowCoord rc{};
  if (p.amode == A_CONV3X3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + (tid >> 3) + 16 * i;
      int mm = m < p.M ? m : 0;
      int ox = mm % p.Wout, t = mm / p.Wout;
      int oy = t % p.Hout, b = t / p.Hout;
      int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
      conv_pix[i] = (b * p.Hin + iy0) * p.Win + ix0;
      conv_yx[i] = m < p.M ? (((iy0 + 64) << 16) | (ix0 + 64)) : -1;
    }
  } else if (p.amode == A_UPCAT) {
    int m = m0 + tid;
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

  auto load_stage = [&](int kb, int s) {
    const uint32_t a_s = base + s * C::kStage;
    const uint32_t b_s = a_s + kABytes;
    if (p.amode == A_UPCAT) {
      const int r = tid, m = m0 + r;
      const int c2 = p.K - p.Cin;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = (kb * 8 + c) * 8;
        const uint32_t dst = a_s + sw128_off(r, c);
        if (m < p.M && k < p.Cin) {
          uint4 a = *reinterpret_cast<const uint4*>(rc.p00 + k);
          uint4 b = *reinterpret_cast<const uint4*>(rc.p01 + k);
          uint4 cc = *reinterpret_cast<const uint4*>(rc.p10 + k);
          uint4 d = *reinterpret_cast<const uint4*>(rc.p11 + k);
          uint4 o = lerp8(a, b, cc, d, rc);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
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
        for (int i = 0; i < 8; ++i) {
          const int r = (tid >> 3) + 16 * i, m = m0 + r;
          bool valid = m < p.M && k < p.K;
          const __nv_bfloat16* src = valid ? p.A + (size_t)m * p.lda + k : p.A;
          cp_async16(a_s + sw128_off(r, c), src, valid);
        }
      } else {  // implicit GEMM over the 9 taps of a dense 3x3 conv: k = tap*Cin + ci
        const int tap = k / p.Cin, ci = k - tap * p.Cin;
        const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (tid >> 3) + 16 * i;
          const int iy = (conv_yx[i] >> 16) - 64 + ky, ix = (conv_yx[i] & 0xFFFF) - 64 + kx;
          bool valid = conv_yx[i] >= 0 && k < p.K && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
          const __nv_bfloat16* src = valid ? p.A + (size_t)(conv_pix[i] + ky * p.Win + kx) * p.Cin + ci : p.A;
          cp_async16(a_s + sw128_off(r, c), src, valid);
        }
      }
    }
    if (tid == 0) {
      mbar_arrive_expect_tx(full_b(s), BN * 128);
      bulk_g2s(b_s, p.W + ((size_t)kb * p.N + n0) * 128, BN * 128, full_b(s));
    }
    cp_async_commit();
  };

  // ---- main loop: S-stage software pipeline ------------------------------------------------------
#pragma unroll
  for (int j = 0; j < S - 1; ++j) {
    if (j < KB) load_stage(j, j);
    else cp_async_commit();
  }
  constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb % S;
    cp_async_wait<S - 2>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(full_b(s), (kb / S) & 1);
      tc_fence_after();
      const uint32_t a_s = base + s * C::kStage;
      const uint64_t adesc = umma_desc_sw128(a_s), bdesc = umma_desc_sw128(a_s + kABytes);
      int ksteps = (p.K - kb * 64 + 15) >> 4;
      ksteps = ksteps > 4 ? 4 : ksteps;
      for (int j = 0; j < ksteps; ++j) umma_bf16(tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (kb | j) != 0);
      umma_commit(mma_done(s));
    }
    const int nk = kb + S - 1;
    if (nk < KB) {
      if (kb >= 1) mbar_wait(mma_done((kb - 1) % S), ((kb - 1) / S) & 1);  // stage being refilled is free
      load_stage(nk, nk % S);
    } else {
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  mbar_wait(mma_done((KB - 1) % S), ((KB - 1) / S) & 1);
  tc_fence_after();

  // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 = output rows m0+32w+lane ---------------------
  const int m = m0 + warp * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t acc[32];
    tmem_ld32(trow + c0, acc);
    tmem_ld_wait();
    if (m < p.M) {
      const int n = n0 + c0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 8 columns per group -> one 16 B store
        float v[8];
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 8 * g));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 8 * g + 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[8 * g + j]) + bb[j];
        if (p.res_pre) {
          const uint4 r = *reinterpret_cast<const uint4*>(p.res_pre + (size_t)m * p.ld_rpre + n + 8 * g);
          const uint32_t* pr = &r.x;
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.rscale + n + 8 * g));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.rscale + n + 8 * g + 4));
          const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] += ss[2 * j] * bf16_lo(pr[j]);
            v[2 * j + 1] += ss[2 * j + 1] * bf16_hi(pr[j]);
          }
        }
        if (p.leaky) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = leaky(v[j]);
        }
        if (p.res_post) {
          const uint4 r = *reinterpret_cast<const uint4*>(p.res_post + (size_t)m * p.ld_rpost + n + 8 * g);
          const uint32_t* pr = &r.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] += bf16_lo(pr[j]);
            v[2 * j + 1] += bf16_hi(pr[j]);
          }
        }
        if (p.post_scale) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.post_scale + n + 8 * g));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.post_scale + n + 8 * g + 4));
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.post_shift + n + 8 * g));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.post_shift + n + 8 * g + 4));
          const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = leaky(ss[j] * v[j] + tt[j]);
        }
        uint4 o;
        o.x = pack_bf16(v[0], v[1]);
        o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]);
        o.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.C + (size_t)m * p.ldc + n + 8 * g) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::kTmemCols);
}

template <int BN, int S>
int launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  dim3 grid(a.N / BN, (a.M + kBM - 1) / kBM);
  gemm_tc_kernel<BN, S><<<grid, 128, Cfg<BN, S>::kSmem, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int BN, int S>
int set_attr() {
  return (int)cudaFuncSetAttribute(gemm_tc_kernel<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg<BN, S>::kSmem);
}

}  // namespace

int gemm_init() {
  int e = 0;
  e |= set_attr<32, 4>();
  e |= set_attr<64, 4>();
  e |= set_attr<128, 3>();
  e |= set_attr<256, 3>();
  return e;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  if (a.N % 32 != 0 || a.K % 8 != 0) return (int)cudaErrorInvalidValue;
  const long mt = (a.M + kBM - 1) / kBM;
  // widest N tile that still yields >= 2 waves of CTAs; otherwise favour more CTAs
  if (a.N % 256 == 0 && mt * (a.N / 256) >= 296) return launch_cfg<256, 3>(a, stream);
  if (a.N % 128 == 0 && (mt * (a.N / 128) >= 148 || a.N % 64 != 0)) return launch_cfg<128, 3>(a, stream);
  if (a.N % 64 == 0 && (mt * (a.N / 64) >= 148 || a.N == 64)) return launch_cfg<64, 4>(a, stream);
  if (a.N % 64 == 0) return launch_cfg<64, 4>(a, stream);
  return launch_cfg<32, 4>(a, stream);
}

}  // namespace casync
