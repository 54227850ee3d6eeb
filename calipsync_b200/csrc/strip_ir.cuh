// Strip-streaming fused InvertedResidual kernel (pw1 -> depthwise 3x3 -> pw2 in one launch); see strip_ir.cu.
#pragma once
#include "common.cuh"

namespace casync {

struct StripArgs {
  const __nv_bfloat16* in;   // block input NHWC [B,W,W,cin]   (decoder: the skip tensor [B,W,W,cin/2])
  const __nv_bfloat16* low;  // decoder only: low-res tensor [B,W/2,W/2,cin/2], bilinearly upsampled on the fly
  __nv_bfloat16* out;        // NHWC [B,W/stride,W/stride,cout], dense
  const uint8_t* W1;         // packed [k-block][2cin rows][128 B]
  const uint8_t* W2;         // packed [k-block][cout rows][128 B]
  const uint8_t* wdp;        // depthwise taps + folded-BN bias as bf16 [2cin/8][10][8]
  int batch;
  unsigned long long* dbg;   // optional per-role cycle counters (developer timing, CASYNC_PHASE_DBG)
  float b1[128];             // folded-BN biases travel as kernel parameters: the drains add them straight from the
  float b2[128];             // constant bank (no shared-memory loads, no registers)
  // strip_tc.cu, last decoder block only: OutConv + outc_bn + sigmoid (module/unet.py:100-106, 342-344) in the epilogue
  void* final_out;           // fp32 NCHW [B,3,160,160] or uint8 HWC [B,160,160,3]; null: store the block's bf16 output
  int final_u8;              // 1: uint8 HWC = floor(p * 255)
  float wo[96], bo[3];       // folded OutConv weights [3][32] and biases
  // strip_tc.cu, `inc` instantiation (InConvDw, module/unet.py:58-67: 6 -> 12 -> 32 at 160x160, fp32 NCHW input)
  const float* x_nchw;       // [B,6,160,160]
  float inc_w1[72];          // [12 hidden][6 cin], BN folded (b1[] above holds the 12 biases)
  float inc_wd[108];         // [9 taps][12 hidden]
  float inc_bd[16];          // 12 biases, zero padded
};

// cin/cout/W/stride/upcat/res select the instantiation.  Returns -1 when the block shape has none, else 0 / cudaError.
int launch_strip_ir(const StripArgs& a, int cin, int cout, int W, int stride, bool upcat, bool res, int num_sms,
                    cudaStream_t st);
bool strip_ir_supported(int cin, int cout, int W, int stride, bool upcat, bool res);
// strip_tc.cu: the same block with the depthwise 3x3 on the tensor cores (64 hidden channels)
int launch_strip_tc(const StripArgs& a, int cin, int cout, int W, int stride, bool upcat, bool res, int num_sms,
                    cudaStream_t st);
bool strip_tc_supported(int cin, int cout, int W, int stride, bool upcat, bool res);
// strip_tc.cu: the input block (x_nchw, inc_*, b1, b2, W2 = the packed pw2 tile, out = x1 NHWC)
int launch_strip_inc(const StripArgs& a, int num_sms, cudaStream_t st);

}  // namespace casync
