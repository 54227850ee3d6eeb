// Fused InvertedResidual kernel (pw1 -> dw3x3 -> pw2 in one launch); see fused_ir.cu.
#pragma once
#include "common.cuh"

namespace casync {

struct FusedArgs {
  const __nv_bfloat16* in;   // block input NHWC [B,W,W,cin]   (decoder: the skip tensor [B,W,W,cin/2])
  const __nv_bfloat16* low;  // decoder only: low-res tensor [B,W/2,W/2,cin/2], bilinearly upsampled on the fly
  __nv_bfloat16* out;        // NHWC [B,W/stride,W/stride,cout] with pixel pitch ldo
  int ldo;
  const uint8_t* W1;         // packed [k-block][2cin rows][128 B]
  const uint8_t* W2;         // packed [k-block][cout rows][128 B]
  const float *wd, *b1, *bd, *b2;
  const uint8_t* wdp;        // weight-streaming instantiations: depthwise taps + bias as bf16 [2cin/8][10][8]
  int W, batch, num_sms;
  int cin, cout, stride;
  bool upcat, res;
  unsigned long long* dbg;   // optional [8] phase-cycle counters (developer timing, CASYNC_PHASE_DBG=1)
};

int launch_fused_ir(const FusedArgs& a, cudaStream_t st);   // -1: shape not instantiated; else 0 / cudaError
bool fused_ir_supported(int cin, int cout, int stride, bool upcat, bool res);

}  // namespace casync
