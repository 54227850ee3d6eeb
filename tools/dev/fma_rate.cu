// Developer microbenchmark: issue rate of HFMA2.BF16 / HFMA2.F16 / FFMA / FFMA2 per SM (independent chains).
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(int iters, float* out, unsigned long long* cyc) {
  float acc = 0;
  long long t0 = clock64();
  if (MODE == 0) {
    __nv_bfloat162 a[8], w = __floats2bfloat162_rn(1.0001f, 0.9999f), c = __floats2bfloat162_rn(0.001f, 0.002f);
    for (int j = 0; j < 8; ++j) a[j] = __floats2bfloat162_rn(threadIdx.x * 0.001f + j, 1.f);
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __hfma2(a[j], w, c);
    for (int j = 0; j < 8; ++j) acc += __low2float(a[j]) + __high2float(a[j]);
  } else if (MODE == 1) {
    __half2 a[8], w = __floats2half2_rn(1.0001f, 0.9999f), c = __floats2half2_rn(0.001f, 0.002f);
    for (int j = 0; j < 8; ++j) a[j] = __floats2half2_rn(threadIdx.x * 0.001f + j, 1.f);
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __hfma2(a[j], w, c);
    for (int j = 0; j < 8; ++j) acc += __low2float(a[j]) + __high2float(a[j]);
  } else if (MODE == 2) {
    float a[8], w = 1.0001f, c = 0.001f;
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 0.001f + j;
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], w, c);
    for (int j = 0; j < 8; ++j) acc += a[j];
  } else {
    unsigned long long a[8], w, c;
    float2 wf = make_float2(1.0001f, 0.9999f), cf = make_float2(0.001f, 0.002f);
    w = *reinterpret_cast<unsigned long long*>(&wf); c = *reinterpret_cast<unsigned long long*>(&cf);
    for (int j = 0; j < 8; ++j) { float2 v = make_float2(threadIdx.x * 0.001f + j, 1.f); a[j] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[j]) : "l"(a[j]), "l"(w), "l"(c));
    for (int j = 0; j < 8; ++j) { float2 v = *reinterpret_cast<float2*>(&a[j]); acc += v.x + v.y; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  float* o; unsigned long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
  const char* names[4] = {"HFMA2.BF16", "HFMA2.F16", "FFMA", "FFMA2(f32x2)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      const int iters = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(iters, o, c);
        if (mode == 1) k<1><<<148, threads>>>(iters, o, c);
        if (mode == 2) k<2><<<148, threads>>>(iters, o, c);
        if (mode == 3) k<3><<<148, threads>>>(iters, o, c);
        cudaDeviceSynchronize();
      }
      unsigned long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      double winstr = (double)iters * 8 * (threads / 32);
      printf("%-14s threads=%4d  %.2f warp-instr/cycle/SM\n", names[mode], threads, winstr / h);
    }
  return 0;
}
