// Developer self-check for the tcgen05 GEMM (not part of the shipped library): random A/W, host-side
// weight packing, all three A modes, compared with a double-precision host loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 dev_gemm_check.cu gemm_tc.cu -o dev_gemm_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_tc.cuh"

using namespace casync;

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

static std::vector<uint8_t> pack_w(const std::vector<float>& w, int N, int K) {
  int KB = (K + 63) / 64;
  std::vector<uint8_t> out((size_t)KB * N * 128, 0);
  for (int kb = 0; kb < KB; ++kb)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < 8; ++c)
        for (int e = 0; e < 8; ++e) {
          int k = kb * 64 + c * 8 + e;
          float v = k < K ? w[(size_t)n * K + k] : 0.f;
          __nv_bfloat16 h = __float2bfloat16(v);
          size_t off = ((size_t)kb * N + n) * 128 + ((c ^ (n & 7)) << 4) + e * 2;
          *reinterpret_cast<__nv_bfloat16*>(&out[off]) = h;
        }
  return out;
}

template <class T>
static T* dev(const std::vector<T>& h) {
  T* d;
  cudaMalloc(&d, h.size() * sizeof(T) + 256);
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

static int check(const char* name, int amode, int B, int Hin, int Win, int Cin, int C2, int N, int stride, int pad,
                 bool extras) {
  int Hout, Wout, K, M;
  if (amode == A_PLAIN) { Hout = Hin; Wout = Win; K = Cin; }
  else if (amode == A_CONV3X3) { Hout = (Hin + 2 * pad - 3) / stride + 1; Wout = (Win + 2 * pad - 3) / stride + 1; K = 9 * Cin; }
  else { Hout = 2 * Hin; Wout = 2 * Win; K = Cin + C2; }
  M = B * Hout * Wout;
  std::vector<float> a((size_t)B * Hin * Win * Cin), a2((size_t)M * (C2 > 0 ? C2 : 1)), w((size_t)N * K), bias(N), rs(N), ps(N), pt(N);
  std::vector<float> rpre((size_t)M * N), rpost((size_t)M * N);
  for (auto& v : a) v = bf(frand());
  for (auto& v : a2) v = bf(frand());
  for (auto& v : w) v = bf(frand() * 0.2f);
  for (auto& v : bias) v = frand();
  for (auto& v : rs) v = 0.5f + frand() * 0.2f;
  for (auto& v : ps) v = 1.f + frand() * 0.2f;
  for (auto& v : pt) v = frand() * 0.1f;
  for (auto& v : rpre) v = bf(frand());
  for (auto& v : rpost) v = bf(frand());
  auto tobf = [](const std::vector<float>& f) { std::vector<__nv_bfloat16> o(f.size()); for (size_t i = 0; i < f.size(); ++i) o[i] = __float2bfloat16(f[i]); return o; };
  auto da = dev(tobf(a)); auto da2 = dev(tobf(a2)); auto dw = dev(pack_w(w, N, K));
  auto db = dev(bias); auto drs = dev(rs); auto dps = dev(ps); auto dpt = dev(pt);
  auto drpre = dev(tobf(rpre)); auto drpost = dev(tobf(rpost));
  std::vector<__nv_bfloat16> hc((size_t)M * N);
  auto dc = dev(hc);
  GemmArgs g{};
  g.amode = amode; g.A = da; g.A2 = da2; g.lda = Cin; g.M = M; g.K = K; g.N = N;
  g.Hin = Hin; g.Win = Win; g.Cin = Cin; g.Hout = Hout; g.Wout = Wout; g.stride = stride; g.pad = pad;
  g.W = dw; g.bias = db; g.leaky = 1; g.C = dc; g.ldc = N;
  if (extras) { g.rscale = drs; g.res_pre = drpre; g.ld_rpre = N; g.res_post = drpost; g.ld_rpost = N; g.post_scale = dps; g.post_shift = dpt; }
  int e = launch_gemm(g, 0);
  cudaError_t ce = cudaDeviceSynchronize();
  if (e || ce) { printf("%s: launch %d sync %s\n", name, e, cudaGetErrorString(ce)); return 1; }
  cudaMemcpy(hc.data(), dc, hc.size() * 2, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  std::vector<float> row(K);
  for (int m = 0; m < M; ++m) {
    int x = m % Wout, y = (m / Wout) % Hout, b = m / (Wout * Hout);
    for (int k = 0; k < K; ++k) {
      float v = 0;
      if (amode == A_PLAIN) v = a[(size_t)m * Cin + k];
      else if (amode == A_CONV3X3) {
        int tap = k / Cin, ci = k % Cin, iy = y * stride - pad + tap / 3, ix = x * stride - pad + tap % 3;
        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) v = a[(((size_t)b * Hin + iy) * Win + ix) * Cin + ci];
      } else if (k < Cin) {
        float sy = (float)(Hin - 1) / (float)(Hout - 1) * y, sx = (float)(Win - 1) / (float)(Wout - 1) * x;
        int y0 = (int)sy, x0 = (int)sx, y1 = y0 + (y0 < Hin - 1), x1 = x0 + (x0 < Win - 1);
        float wy1 = sy - y0, wy0 = 1 - wy1, wx1 = sx - x0, wx0 = 1 - wx1;
        auto at = [&](int yy, int xx) { return a[(((size_t)b * Hin + yy) * Win + xx) * Cin + k]; };
        v = bf(wy0 * (wx0 * at(y0, x0) + wx1 * at(y0, x1)) + wy1 * (wx0 * at(y1, x0) + wx1 * at(y1, x1)));
      } else v = a2[(size_t)m * C2 + (k - Cin)];
      row[k] = v;
    }
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)row[k] * w[(size_t)n * K + k];
      double v = acc + bias[n];
      if (extras) v += rs[n] * rpre[(size_t)m * N + n];
      v = v > 0 ? v : 0.01 * v;
      if (extras) { v += rpost[(size_t)m * N + n]; v = ps[n] * v + pt[n]; v = v > 0 ? v : 0.01 * v; }
      double got = __bfloat162float(hc[(size_t)m * N + n]);
      maxerr = fmax(maxerr, fabs(got - v));
      maxref = fmax(maxref, fabs(v));
    }
  }
  int bad = !(maxerr <= 0.01 * maxref + 1e-3);
  printf("%-28s M=%d K=%d N=%d maxerr=%.5f maxref=%.3f %s\n", name, M, K, N, maxerr, maxref, bad ? "FAIL" : "ok");
  return bad;
}

int main() {
  if (gemm_init()) { printf("gemm_init failed\n"); return 2; }
  int bad = 0;
  bad += check("plain K64 N64", A_PLAIN, 1, 10, 13, 64, 0, 64, 1, 0, false);
  bad += check("plain K32 N64", A_PLAIN, 1, 16, 16, 32, 0, 64, 1, 0, false);
  bad += check("plain K128 N32 extras", A_PLAIN, 2, 20, 20, 128, 0, 32, 1, 0, true);
  bad += check("plain K512 N128", A_PLAIN, 3, 10, 10, 512, 0, 128, 1, 0, true);
  bad += check("plain K1024 N256 big", A_PLAIN, 400, 10, 10, 1024, 0, 512, 1, 0, false);
  bad += check("plain K2048 N1024", A_PLAIN, 2, 10, 10, 2048, 0, 1024, 1, 0, true);
  bad += check("conv3x3 s2 p1", A_CONV3X3, 2, 32, 32, 128, 0, 256, 2, 1, false);
  bad += check("conv3x3 s2 p3", A_CONV3X3, 2, 16, 16, 256, 0, 512, 2, 3, true);
  bad += check("upcat 32+32", A_UPCAT, 1, 20, 20, 32, 32, 128, 1, 0, false);
  bad += check("upcat 256+256", A_UPCAT, 2, 10, 10, 256, 256, 1024, 1, 0, true);
  printf(bad ? "DEV GEMM CHECK: %d FAILED\n" : "DEV GEMM CHECK: all ok\n", bad);
  return bad;
}
