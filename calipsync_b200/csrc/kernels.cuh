// CUDA-core kernels of the CASync generator forward: everything that is not a dense contraction.
#pragma once
#include "common.cuh"

namespace casync {

// `inc` (InConvDw, module/unet.py:58-67): one InvertedResidual 6 -> (12) -> 32 at 160x160 in one kernel: pw1 and the
// depthwise conv on CUDA cores (K = 6 is far below a UMMA tile), pw2 (12 -> 32, K padded to 16) on the tensor core.
// The small weights travel as kernel parameters (constant bank); pw2 additionally as a packed bf16 UMMA tile.
struct IncParams {
  float w1[12 * 6];   // [hidden][cin], BN folded
  float b1[12];
  float wd[9 * 12];   // [tap][hidden]
  float bd[12];
  float w2[32 * 12];  // [cout][hidden]
  float b2[32];
};
struct OutcParams {   // OutConv + outc_bn folded (module/unet.py:100-106, 342-343)
  float w[3 * 32];
  float b[3];
  float pad_;
};

int launch_inc(const float* x_nchw, __nv_bfloat16* x1_nhwc, const uint8_t* w2_tile, const IncParams& w, int batch,
               cudaStream_t st);
// depthwise 3x3, pad 1, stride 1|2, + folded BN bias + LeakyReLU.  NHWC bf16 -> NHWC bf16; wdp: taps + bias as bf16,
// [C/8][10][8] (9 taps, then the bias, per 8-channel chunk).
int launch_dw3x3(const __nv_bfloat16* in, __nv_bfloat16* out, const uint8_t* wdp, int batch, int H, int W, int C,
                 int stride, cudaStream_t st);
// caller-side input assembly on the device: uint8 HWC crops -> x [B,6,160,160] (reference face + masked face, /255)
// and HuBERT features [T,2,1024] + frame indices -> audio windows [B,32,32,32] (infer_api.py:99-145, 238-245)
int launch_prepare_inputs(const uint8_t* crops, const float* feats, int T, const int* frame_idx, float* x, float* audio,
                          int batch, cudaStream_t st);
// audio window fp32 [B,32,32,32] NCHW -> bf16 NHWC
int launch_audio_prep(const float* audio, __nv_bfloat16* out, int batch, cudaStream_t st);
// softmax(q k^T) v core of CrossAttention (module/unet.py:209-217) for one attention block, on the tensor cores:
//   out[m, c] = gamma * sum_j softmax_j(q[m,:].k[j,:]) v[j,c] + x[m,c]   (per frame: 100 queries x 100 keys)
// q [M, 64] (pitch ldq), k [M, 64] (pitch ldk) row-major; vt = V^T per frame: [B][4 blocks][512][128 keys] with the
// block offset already applied (frame stride 4*512*128); x, out: [M, 512] with pitches ldx, 512.
int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     const __nv_bfloat16* x, int ldx, __nv_bfloat16* out, float gamma, int batch, cudaStream_t st);
// kx = leaky(bn_kx(tx + ox0 + ox1 + ox2 + ox3))  (module/unet.py:329-336), fp32 sum of the bf16 stage tensors
int launch_sum5(const __nv_bfloat16* tx, const __nv_bfloat16* o0, const __nv_bfloat16* o1, const __nv_bfloat16* o2,
                const __nv_bfloat16* o3, const float* s, const float* t, __nv_bfloat16* kx, long rows,
                cudaStream_t st);
// sigmoid(outc_bn(outc(x)))  ->  fp32 NCHW [B,3,160,160] or uint8 HWC floor(p*255)
int launch_outc(const __nv_bfloat16* x, void* out, const OutcParams& w, int batch, int u8_hwc, cudaStream_t st);
// paste-back blend (infer_api.py:333-346): frames[region] = uint8(crop * m + frames[region] * (1 - m)) in float64
// cat([bilinear x2 (align_corners) of low [B,H/2,W/2,C2], skip [B,H,W,C2]]) -> out [B,H,W,2*C2], NHWC bf16: the decoder's
// upsample + concat (module/unet.py:82-97) materialised, same arithmetic as the GEMM's gathered A producer
int launch_upcat(const __nv_bfloat16* low, const __nv_bfloat16* skip, __nv_bfloat16* out, int batch, int H, int C2,
                 cudaStream_t st);
int launch_blend_paste(uint8_t* frames, int H, int W, const uint8_t* crops, int ldc, const uint8_t* face, const float* soft,
                       const int* rects, int batch, cudaStream_t st);
int kernels_init();

}  // namespace casync
