// CUDA-core kernels (sm_100a): fused `inc` block, depthwise 3x3, audio layout change, attention core,
// stage sum + BN, output head.  All activations NHWC bf16, all arithmetic fp32.
#include "kernels.cuh"

#include <cuda_bf16.h>

namespace casync {

namespace {

// ------------------------------------------------------------------------------------------------
// inc: 32x8 output tile per CTA, 34x10 input halo.  The depthwise conv zero-pads the HIDDEN tensor
// (module/unet.py:21-27), so hidden values of out-of-image halo pixels are exactly 0, not pw1(0).
// ------------------------------------------------------------------------------------------------
constexpr int kIncTW = 32, kIncTH = 8, kIncHW = kIncTW + 2, kIncHH = kIncTH + 2, kIncHalo = kIncHW * kIncHH;

__global__ void __launch_bounds__(256) inc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                  const __grid_constant__ IncParams w) {
  __shared__ float s_in[6][kIncHalo];
  __shared__ float s_h[12][kIncHalo];
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int x0 = blockIdx.x * kIncTW, y0 = blockIdx.y * kIncTH, b = blockIdx.z;
  const float* xb = x + (size_t)b * 6 * 25600;
  for (int idx = tid; idx < 6 * kIncHalo; idx += 256) {
    int c = idx / kIncHalo, r = idx - c * kIncHalo;
    int yy = r / kIncHW, xx = r - yy * kIncHW;
    int gy = y0 - 1 + yy, gx = x0 - 1 + xx;
    float v = 0.f;
    if (gy >= 0 && gy < 160 && gx >= 0 && gx < 160) v = __ldg(xb + (size_t)c * 25600 + gy * 160 + gx);
    s_in[c][r] = v;
  }
  __syncthreads();
  for (int r = tid; r < kIncHalo; r += 256) {
    int yy = r / kIncHW, xx = r - yy * kIncHW;
    int gy = y0 - 1 + yy, gx = x0 - 1 + xx;
    bool inside = gy >= 0 && gy < 160 && gx >= 0 && gx < 160;
    float in[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) in[c] = s_in[c][r];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      float a = w.b1[j];
#pragma unroll
      for (int c = 0; c < 6; ++c) a = fmaf(w.w1[j * 6 + c], in[c], a);
      s_h[j][r] = inside ? leaky(a) : 0.f;
    }
  }
  __syncthreads();
  const int ty = tid >> 5, tx = tid & 31;
  float h2[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float a = w.bd[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) a = fmaf(w.wd[(ky * 3 + kx) * 12 + j], s_h[j][(ty + ky) * kIncHW + tx + kx], a);
    h2[j] = leaky(a);
  }
  __nv_bfloat16* o = out + ((size_t)b * 25600 + (size_t)(y0 + ty) * 160 + x0 + tx) * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int n = g * 8 + e;
      float a = w.b2[n];
#pragma unroll
      for (int j = 0; j < 12; ++j) a = fmaf(w.w2[n * 12 + j], h2[j], a);
      v[e] = leaky(a);
    }
    uint4 q;
    q.x = pack_bf16(v[0], v[1]);
    q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]);
    q.w = pack_bf16(v[6], v[7]);
    reinterpret_cast<uint4*>(o)[g] = q;
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 (+ folded BN bias + LeakyReLU) for the blocks that are not fully fused.  CTA = (64-channel
// slice, band of output rows, frame).  The input band with a zero border -- (rows*S+2) x (W+2) pixels x 128 B --
// is fetched with cp.async, every load of the CTA in flight at once (the old register-window version had
// ~16 KB in flight per SM and ran at 1.3 TB/s); then unit = (output pixel, 16-byte channel chunk): nine
// conflict-free LDS.128, packed bf16x2 FMAs (same arithmetic as the fused kernel's depthwise), one 16 B store.
// ------------------------------------------------------------------------------------------------
constexpr int kDwSmemMax = 56 * 1024;

template <int STRIDE>
__global__ void __launch_bounds__(256) dw3x3_kernel(const __nv_bfloat16* __restrict__ in,
                                                    __nv_bfloat16* __restrict__ out, const float* __restrict__ wd,
                                                    const float* __restrict__ bd, int H, int W, int C, int Ho,
                                                    int Wo, int BR) {
  extern __shared__ uint4 dw_tile[];   // [rows_in][W + 2][8 chunks of 8 channels]
  pdl_launch_dependents();
  const int tid = threadIdx.x, chunk = tid & 7;
  const int c0 = blockIdx.x * 64, oy0 = blockIdx.y * BR, b = blockIdx.z;
  const int rows_out = BR < Ho - oy0 ? BR : Ho - oy0;
  const int rows_in = (rows_out - 1) * STRIDE + 3, TW = W + 2;
  const int iy0 = oy0 * STRIDE - 1;
  // taps / bias of this thread's 8 channels (constants: fetched before waiting for the previous kernel)
  const int c = c0 + chunk * 8;
  __nv_bfloat162 wt[9][4], wb[4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wd + t * C + c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wd + t * C + c + 4));
    wt[t][0] = __floats2bfloat162_rn(w0.x, w0.y);
    wt[t][1] = __floats2bfloat162_rn(w0.z, w0.w);
    wt[t][2] = __floats2bfloat162_rn(w1.x, w1.y);
    wt[t][3] = __floats2bfloat162_rn(w1.z, w1.w);
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bd + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bd + c + 4));
    wb[0] = __floats2bfloat162_rn(b0.x, b0.y);
    wb[1] = __floats2bfloat162_rn(b0.z, b0.w);
    wb[2] = __floats2bfloat162_rn(b1.x, b1.y);
    wb[3] = __floats2bfloat162_rn(b1.z, b1.w);
  }
  const __nv_bfloat162 kslope = __floats2bfloat162_rn(kLeaky, kLeaky);
  pdl_wait();   // the hidden tensor is the previous kernel's output
  const __nv_bfloat16* ib = in + (size_t)b * H * W * C + c;
  const uint32_t tile_s = smem_u32(dw_tile);
  const int npx = rows_in * TW;
  for (int px = tid >> 3; px < npx; px += 32) {      // 256 % 8 == 0: a thread always copies its own chunk column
    const int ty = px / TW, tx = px - ty * TW;
    const int iy = iy0 + ty, ix = tx - 1;
    const bool valid = iy >= 0 && iy < H && ix >= 0 && ix < W;
    cp_async16(tile_s + (uint32_t)(px * 8 + chunk) * 16u, valid ? ib + ((size_t)iy * W + ix) * C : ib, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int nunits = rows_out * Wo;
  __nv_bfloat16* ob = out + ((size_t)b * Ho + oy0) * Wo * C + c;
  for (int p = tid >> 3; p < nunits; p += 32) {
    const int oy = p / Wo, ox = p - oy * Wo;
    const uint4* t0 = dw_tile + ((oy * STRIDE) * TW + ox * STRIDE) * 8 + chunk;
    __nv_bfloat162 a[4] = {wb[0], wb[1], wb[2], wb[3]};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint4 v = t0[(ky * TW + kx) * 8];
        const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = __hfma2(wt[ky * 3 + kx][q], pv[q], a[q]);
      }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) po[q] = __hmax2(a[q], __hmul2(a[q], kslope));
    *reinterpret_cast<uint4*>(ob + (size_t)p * C) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// audio window: fp32 [B,32(c),32,32] -> bf16 [B,32,32,32(c)]; one thread = one pixel (reads are coalesced
// across the warp for every channel, the 64 B result is written with four 16 B stores).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) audio_prep_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ out,
                                                         long npix) {
  pdl_launch_dependents();
  pdl_wait();
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const long b = p >> 10, r = p & 1023;
  const float* src = a + b * 32768 + r;
  uint32_t w[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) w[c] = pack_bf16(__ldg(src + (2 * c) * 1024), __ldg(src + (2 * c + 1) * 1024));
  uint4* dst = reinterpret_cast<uint4*>(out + p * 32);
#pragma unroll
  for (int g = 0; g < 4; ++g) dst[g] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
}

// ------------------------------------------------------------------------------------------------
// attention core.  CTA = (channel quarter, frame), 10 warps; warp w owns query rows 10w..10w+9 as two
// groups of 5.  S = q k^T and the row softmax stay in registers (lane owns keys lane+32t), P goes to
// shared memory only for the warp's own rows, then O = P V over this CTA's 128 value channels.
// Softmax has no 1/sqrt(d) scale (module/unet.py:212-213).
// ------------------------------------------------------------------------------------------------
constexpr int kT = 100;  // tokens per frame (10x10)
struct AttnSmem {
  uint32_t q2[kT][32];     // q[i][2d..2d+1] packed bf16x2
  uint32_t kT2[32][kT];    // k[j][2d..2d+1] packed, transposed
  uint2 v4[kT][32];        // v[j][4*lane..4*lane+3] (this CTA's 128 channels)
  float P[kT][kT];
};

__global__ void __launch_bounds__(320) attention_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                        const __nv_bfloat16* __restrict__ k,
                                                        const __nv_bfloat16* __restrict__ v, int ldkv,
                                                        const __nv_bfloat16* __restrict__ x, int ldx,
                                                        __nv_bfloat16* __restrict__ out, float gamma) {
  extern __shared__ uint8_t smem_raw[];
  AttnSmem& s = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  pdl_wait();
  const int quarter = blockIdx.x;
  const size_t row0 = (size_t)blockIdx.y * kT;
  for (int idx = tid; idx < kT * 32; idx += 320) {
    const int j = idx >> 5, d2 = idx & 31;
    s.q2[j][d2] = __ldg(reinterpret_cast<const uint32_t*>(q + (row0 + j) * ldq) + d2);
    s.kT2[d2][j] = __ldg(reinterpret_cast<const uint32_t*>(k + (row0 + j) * ldkv) + d2);
    s.v4[j][d2] = __ldg(reinterpret_cast<const uint2*>(v + (row0 + j) * ldkv + quarter * 128) + d2);
  }
  __syncthreads();
#pragma unroll 1
  for (int grp = 0; grp < 2; ++grp) {
    const int i0 = warp * 10 + grp * 5;
    float sc[5][4];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int t = 0; t < 4; ++t) sc[r][t] = 0.f;
#pragma unroll 4
    for (int d2 = 0; d2 < 32; ++d2) {
      uint32_t kk[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) kk[t] = (lane + 32 * t < kT) ? s.kT2[d2][lane + 32 * t] : 0u;
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const uint32_t qq = s.q2[i0 + r][d2];
        const float ql = bf16_lo(qq), qh = bf16_hi(qq);
#pragma unroll
        for (int t = 0; t < 4; ++t) sc[r][t] = fmaf(qh, bf16_hi(kk[t]), fmaf(ql, bf16_lo(kk[t]), sc[r][t]));
      }
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < kT) mx = fmaxf(mx, sc[r][t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        sc[r][t] = (lane + 32 * t < kT) ? __expf(sc[r][t] - mx) : 0.f;
        sum += sc[r][t];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < kT) s.P[i0 + r][lane + 32 * t] = sc[r][t] * inv;
    }
    __syncwarp();
    float acc[5][4];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int j = 0; j < kT; j += 4) {
      float4 pr[5];
#pragma unroll
      for (int r = 0; r < 5; ++r) pr[r] = *reinterpret_cast<const float4*>(&s.P[i0 + r][j]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint2 vv = s.v4[j + jj][lane];
        const float v0 = bf16_lo(vv.x), v1 = bf16_hi(vv.x), v2 = bf16_lo(vv.y), v3 = bf16_hi(vv.y);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const float pp = jj == 0 ? pr[r].x : jj == 1 ? pr[r].y : jj == 2 ? pr[r].z : pr[r].w;
          acc[r][0] = fmaf(pp, v0, acc[r][0]);
          acc[r][1] = fmaf(pp, v1, acc[r][1]);
          acc[r][2] = fmaf(pp, v2, acc[r][2]);
          acc[r][3] = fmaf(pp, v3, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const size_t off = (row0 + i0 + r) * 512 + quarter * 128 + lane * 4;
      const uint2 xx = __ldg(reinterpret_cast<const uint2*>(x + (row0 + i0 + r) * ldx + quarter * 128 + lane * 4));
      uint2 o;
      o.x = pack_bf16(fmaf(gamma, acc[r][0], bf16_lo(xx.x)), fmaf(gamma, acc[r][1], bf16_hi(xx.x)));
      o.y = pack_bf16(fmaf(gamma, acc[r][2], bf16_lo(xx.y)), fmaf(gamma, acc[r][3], bf16_hi(xx.y)));
      *reinterpret_cast<uint2*>(out + off) = o;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sum5_kernel(const uint4* __restrict__ tx, const uint4* __restrict__ o0,
                                                   const uint4* __restrict__ o1, const uint4* __restrict__ o2,
                                                   const uint4* __restrict__ o3, const float* __restrict__ s,
                                                   const float* __restrict__ t, uint4* __restrict__ kx, long n8) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int c = (int)(i & 127) * 8;  // 1024 channels = 128 chunks of 8
  const uint4 a = __ldg(tx + i), b = __ldg(o0 + i), cc = __ldg(o1 + i), d = __ldg(o2 + i), e = __ldg(o3 + i);
  const uint32_t *pa = &a.x, *pb = &b.x, *pc = &cc.x, *pd = &d.x, *pe = &e.x;
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float lo = bf16_lo(pa[j]) + bf16_lo(pb[j]) + bf16_lo(pc[j]) + bf16_lo(pd[j]) + bf16_lo(pe[j]);
    float hi = bf16_hi(pa[j]) + bf16_hi(pb[j]) + bf16_hi(pc[j]) + bf16_hi(pd[j]) + bf16_hi(pe[j]);
    lo = leaky(fmaf(__ldg(s + c + 2 * j), lo, __ldg(t + c + 2 * j)));
    hi = leaky(fmaf(__ldg(s + c + 2 * j + 1), hi, __ldg(t + c + 2 * j + 1)));
    po[j] = pack_bf16(lo, hi);
  }
  kx[i] = o;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) outc_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ out,
                                                   const __grid_constant__ OutcParams w, long npix, int u8_hwc) {
  pdl_launch_dependents();
  pdl_wait();
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + p * 32);
  float a0 = w.b[0], a1 = w.b[1], a2 = w.b[2];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 v = __ldg(src + g);
    const uint32_t* pv = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16_lo(pv[j]), hi = bf16_hi(pv[j]);
      const int c = g * 8 + 2 * j;
      a0 = fmaf(w.w[c + 1], hi, fmaf(w.w[c], lo, a0));
      a1 = fmaf(w.w[32 + c + 1], hi, fmaf(w.w[32 + c], lo, a1));
      a2 = fmaf(w.w[64 + c + 1], hi, fmaf(w.w[64 + c], lo, a2));
    }
  }
  const float s0 = 1.f / (1.f + __expf(-a0)), s1 = 1.f / (1.f + __expf(-a1)), s2 = 1.f / (1.f + __expf(-a2));
  if (u8_hwc) {  // floor(p*255) like `np.array(pred*255, dtype=np.uint8)` (infer_api.py:265-266)
    uint8_t* o = reinterpret_cast<uint8_t*>(out) + p * 3;
    o[0] = (uint8_t)(s0 * 255.f);
    o[1] = (uint8_t)(s1 * 255.f);
    o[2] = (uint8_t)(s2 * 255.f);
  } else {
    const long b = p / 25600, r = p - b * 25600;
    float* o = reinterpret_cast<float*>(out) + b * 76800 + r;
    o[0] = s0;
    o[25600] = s1;
    o[51200] = s2;
  }
}

}  // namespace

int kernels_init() {
  int e = (int)cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmem));
  e |= (int)cudaFuncSetAttribute(dw3x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemMax);
  e |= (int)cudaFuncSetAttribute(dw3x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemMax);
  return e;
}

int launch_inc(const float* x, __nv_bfloat16* out, const IncParams& w, int batch, cudaStream_t st) {
  dim3 grid(160 / kIncTW, 160 / kIncTH, batch);
  return (int)launch_pdl(inc_kernel, grid, dim3(256), 0, st, x, out, w);
}

int launch_dw3x3(const __nv_bfloat16* in, __nv_bfloat16* out, const float* wd, const float* bd, int batch, int H,
                 int W, int C, int stride, cudaStream_t st) {
  if (C % 64 != 0 || (stride != 1 && stride != 2)) return (int)cudaErrorInvalidValue;
  const int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? W / 2 : W;
  // rows per band: the input band (BR*stride + 2 rows of (W+2) pixels x 128 B) must fit the smem budget
  const int row_bytes = (W + 2) * 128;
  int BR = ((kDwSmemMax - 1024) / row_bytes - 3) / stride + 1;
  if (BR < 1) return (int)cudaErrorInvalidValue;
  if (BR > Ho) BR = Ho;
  const int bands = (Ho + BR - 1) / BR;
  BR = (Ho + bands - 1) / bands;   // balance the bands
  const size_t smem = (size_t)((BR - 1) * stride + 3) * row_bytes;
  const dim3 grid((unsigned)(C / 64), (unsigned)bands, (unsigned)batch);
  if (stride == 2) return (int)launch_pdl(dw3x3_kernel<2>, grid, dim3(256), smem, st, in, out, wd, bd, H, W, C, Ho, Wo, BR);
  return (int)launch_pdl(dw3x3_kernel<1>, grid, dim3(256), smem, st, in, out, wd, bd, H, W, C, Ho, Wo, BR);
}

int launch_audio_prep(const float* audio, __nv_bfloat16* out, int batch, cudaStream_t st) {
  const long npix = (long)batch * 1024;
  return (int)launch_pdl(audio_prep_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, st, audio, out, npix);
}

int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, const __nv_bfloat16* v, int ldkv,
                     const __nv_bfloat16* x, int ldx, __nv_bfloat16* out, float gamma, int batch, cudaStream_t st) {
  return (int)launch_pdl(attention_kernel, dim3(4, batch), dim3(320), sizeof(AttnSmem), st, q, ldq, k, v, ldkv, x, ldx,
                         out, gamma);
}

int launch_sum5(const __nv_bfloat16* tx, const __nv_bfloat16* o0, const __nv_bfloat16* o1, const __nv_bfloat16* o2,
                const __nv_bfloat16* o3, const float* s, const float* t, __nv_bfloat16* kx, long rows,
                cudaStream_t st) {
  const long n8 = rows * 128;
  return (int)launch_pdl(sum5_kernel, dim3((unsigned)((n8 + 255) / 256)), dim3(256), 0, st,
                         reinterpret_cast<const uint4*>(tx), reinterpret_cast<const uint4*>(o0),
                         reinterpret_cast<const uint4*>(o1), reinterpret_cast<const uint4*>(o2),
                         reinterpret_cast<const uint4*>(o3), s, t, reinterpret_cast<uint4*>(kx), n8);
}

int launch_outc(const __nv_bfloat16* x, void* out, const OutcParams& w, int batch, int u8_hwc, cudaStream_t st) {
  const long npix = (long)batch * 25600;
  return (int)launch_pdl(outc_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, st, x, out, w, npix, u8_hwc);
}

}  // namespace casync
