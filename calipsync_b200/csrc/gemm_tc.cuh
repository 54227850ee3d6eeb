// tcgen05 GEMM for every dense contraction of the CASync generator (1x1 convs, Linear, dense 3x3 convs
// as implicit GEMM, the decoder's first 1x1 with bilinear-upsample + skip-concat fused into the A
// producer).  C[M,N] = epilogue(A[M,K] . W[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"

namespace casync {

enum AMode : int { A_PLAIN = 0, A_CONV3X3 = 1, A_UPCAT = 2 };

struct GemmArgs {
  int amode;
  const __nv_bfloat16* A;   // PLAIN: [M, lda]; CONV3X3: NHWC [B,Hin,Win,Cin]; UPCAT: low-res NHWC [B,Hin,Win,Cin]
  const __nv_bfloat16* A2;  // UPCAT: skip NHWC [B,Hout,Wout,K-Cin]
  int lda;
  int M, K, N;
  int Hin, Win, Cin, Hout, Wout, stride, pad;
  const uint8_t* W;         // packed: k-block kb, row n -> 128 B (64 bf16, SWIZZLE_128B image) at (kb*N+n)*128
  const float* bias;        // [N]
  const float* rscale;      // [N] or null: v += rscale[n] * res_pre[m,n]   (before the activation)
  const __nv_bfloat16* res_pre;
  int ld_rpre;
  int leaky;                // LeakyReLU(0.01) after bias/res_pre
  const __nv_bfloat16* res_post;  // added after the activation (InvertedResidual skip, module/unet.py:38)
  int ld_rpost;
  const float* post_scale;  // optional trailing BN + LeakyReLU (audio bn7, module/unet.py:193)
  const float* post_shift;
  __nv_bfloat16* C;
  int ldc;
  int max_ctas;             // 0: one CTA per SM; >0: cap (two branches of the forward sharing the GPU on two streams)
  // key/value GEMM of the attention blocks: columns >= vt_col0 (the four 512-channel value projections) are stored
  // TRANSPOSED per frame, vt[frame][j][channel][key] with 128 keys per row (100 used), so that the attention kernel
  // can use V^T as a K-major B operand of tcgen05.mma (module/unet.py:215: out = V . attn^T)
  __nv_bfloat16* vt;
  int vt_col0;
  // first 1x1 conv of an InvertedResidual on a dw_w x dw_w image with dw_w^2 <= 100 pixels: apply the block's depthwise
  // 3x3 (+ folded-BN bias + LeakyReLU; taps dwp = bf16 [N/8][10][8]) in the epilogue and store ITS output to C [M, N]
  int dw_epi, dw_w;
  const uint8_t* dwp;
  unsigned long long* dbg;  // developer timing (CASYNC_GEMM_DBG=<label substring>): per-role cycle counters [16]
};

int launch_gemm(const GemmArgs& a, cudaStream_t stream);  // returns 0 or cudaError
int gemm_init();                                          // set smem attributes (once per process)
// 2-D TMA tensor map over a row-major bf16 matrix [rows, cols] with pitch ld (elements): box = 64 columns x box_rows
// rows, SWIZZLE_128B or unswizzled.  tm_out: 128-byte CUtensorMap, 64-byte aligned.  Needs gemm_init().
// SMs a launch may count on when it chooses its tile shape (0 = all): set around work that shares the GPU with another lane
void gemm_set_cost_cap(int ctas);
int gemm_encode_map(void* tm_out, const __nv_bfloat16* A, int rows, int cols, int ld, int box_rows, bool swizzle128);
int gemm_num_sms();

}  // namespace casync
