// Developer harness: drives libcasync_b200.so through the C ABI without Python (starts in ~1 s, so an ncu
// pass costs seconds of box time instead of minutes).  Random weights in the library's packed schema -- timing
// and profiling only, parity is tested from tests/ through the same ABI.
//   casync_run [batch=64] [iters=20] [profile=0|1] [u8=0|1]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/casync_b200.h"

static uint32_t rng_state = 12345u;
static float frand() {  // uniform in [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                    \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  const int profile = argc > 3 ? atoi(argv[3]) : 0;
  const unsigned flags = (argc > 4 && atoi(argv[4])) ? CASYNC_F_OUT_U8_HWC : CASYNC_F_BF16;

  const int n = casync_weight_entry_count();
  std::vector<int64_t> off(n);
  size_t total = 0;
  std::vector<std::string> names(n);
  std::vector<size_t> sizes(n);
  for (int i = 0; i < n; ++i) {
    const char* nm;
    size_t by;
    casync_weight_entry(i, &nm, &by);
    names[i] = nm;
    sizes[i] = by;
    off[i] = (int64_t)total;
    total = (total + by + 255) & ~(size_t)255;
  }
  std::vector<uint8_t> blob(total, 0);
  for (int i = 0; i < n; ++i) {
    const std::string part = names[i].substr(names[i].find('|') + 1);
    const bool is_bf16 = part == "w1" || part == "w2" || part == "w" || part == "kv_w" || part == "p1q_w" || part == "b1_w" || part == "w2t" || part == "wdp";
    if (is_bf16) {
      uint16_t* p = reinterpret_cast<uint16_t*>(blob.data() + off[i]);
      for (size_t j = 0; j < sizes[i] / 2; ++j) p[j] = f2bf(0.03f * frand());
    } else {
      float* p = reinterpret_cast<float*>(blob.data() + off[i]);
      const float scale = (part == "rs" || part == "b1_rs" || part == "s") ? 1.0f : 0.1f;
      for (size_t j = 0; j < sizes[i] / 4; ++j) p[j] = scale == 1.0f ? 1.0f + 0.1f * frand() : scale * frand();
    }
  }
  // the packed bf16 depthwise taps ("wdp", [hid/8][10][8]) must agree with the fp32 taps / bias ("wd" [9][hid], "bd")
  for (int i = 0; i < n; ++i) {
    const size_t bar = names[i].find('|');
    if (names[i].substr(bar + 1) != "wdp") continue;
    const std::string pre = names[i].substr(0, bar + 1);
    int iwd = -1, ibd = -1;
    for (int k = 0; k < n; ++k) {
      if (names[k] == pre + "wd") iwd = k;
      if (names[k] == pre + "bd") ibd = k;
    }
    const int hid = (int)(sizes[i] / 20);
    const float* wd = reinterpret_cast<const float*>(blob.data() + off[iwd]);
    const float* bd = reinterpret_cast<const float*>(blob.data() + off[ibd]);
    uint16_t* o = reinterpret_cast<uint16_t*>(blob.data() + off[i]);
    for (int c = 0; c < hid / 8; ++c)
      for (int t = 0; t < 10; ++t)
        for (int e = 0; e < 8; ++e) o[(c * 10 + t) * 8 + e] = f2bf(t < 9 ? wd[t * hid + c * 8 + e] : bd[c * 8 + e]);
  }
  void* dblob;
  CK(cudaMalloc(&dblob, total));
  CK(cudaMemcpy(dblob, blob.data(), total, cudaMemcpyHostToDevice));
  casync_plan* plan = nullptr;
  if (casync_plan_create(blob.data(), dblob, total, off.data(), n, &plan)) {
    fprintf(stderr, "plan_create: %s\n", casync_last_error());
    return 1;
  }
  const size_t wsb = casync_workspace_bytes(plan, B);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  const int nsets = 3;
  std::vector<float*> xs(nsets), as(nsets);
  const size_t xn = (size_t)B * 6 * 25600, an = (size_t)B * 32768;
  {
    std::vector<float> hx(xn), ha(an);
    for (int s = 0; s < nsets; ++s) {
      for (auto& v : hx) v = 0.5f + 0.5f * frand();
      for (auto& v : ha) v = frand();
      CK(cudaMalloc(&xs[s], xn * 4));
      CK(cudaMalloc(&as[s], an * 4));
      CK(cudaMemcpy(xs[s], hx.data(), xn * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(as[s], ha.data(), an * 4, cudaMemcpyHostToDevice));
    }
  }
  void* out;
  CK(cudaMalloc(&out, (size_t)B * 76800 * 4));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int i = 0; i < 3 * nsets; ++i)   // every input set three times: eager, graph capture, first replay
    if (casync_forward(plan, xs[i % nsets], as[i % nsets], out, ws, B, flags, st)) {
      fprintf(stderr, "forward: %s\n", casync_last_error());
      return 1;
    }
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < iters; ++i) casync_forward(plan, xs[i % nsets], as[i % nsets], out, ws, B, flags, st);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  cudaEventRecord(e1, st);
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double host_ms = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / iters;
  printf("batch %d: %.4f ms/forward -> %.0f frames/s  (%lld launches/forward, host enqueue %.3f ms/forward, workspace %.1f MB)\n",
         B, ms / iters, B * iters / (ms * 1e-3), (long long)casync_launches_per_forward(plan, B), host_ms, wsb / 1e6);
  {   // FNV-1a of the last output: lets two runs (e.g. CASYNC_NO_CHAIN=0|1) be compared bit for bit
    const size_t ob = (size_t)B * 76800 * ((flags & CASYNC_F_OUT_U8_HWC) ? 1 : 4);
    std::vector<uint8_t> ho(ob);
    CK(cudaMemcpy(ho.data(), out, ob, cudaMemcpyDeviceToHost));
    uint64_t h = 1469598103934665603ull;
    double sum = 0;
    for (size_t i = 0; i < ob; ++i) h = (h ^ ho[i]) * 1099511628211ull;
    if (!(flags & CASYNC_F_OUT_U8_HWC))
      for (size_t i = 0; i < ob / 4; ++i) sum += reinterpret_cast<const float*>(ho.data())[i];
    printf("output hash %016llx  mean %.6f\n", (unsigned long long)h, sum / (ob / 4.0));
    if (const char* dump = getenv("CASYNC_DUMP_WS")) {   // whole workspace (stage buffers), for bisecting a mismatch
      std::vector<uint8_t> hw(wsb);
      CK(cudaMemcpy(hw.data(), ws, wsb, cudaMemcpyDeviceToHost));
      FILE* f = fopen(dump, "wb");
      if (f) {
        fwrite(hw.data(), 1, wsb, f);
        fclose(f);
      }
    }
    if (const char* dump = getenv("CASYNC_DUMP")) {
      FILE* f = fopen(dump, "wb");
      if (f) {
        fwrite(ho.data(), 1, ob, f);
        fclose(f);
      }
    }
  }
  if (profile) {
    std::vector<casync_launch_record> recs(1024);
    std::vector<double> acc;
    int nr = 0;
    const int reps = 5;
    for (int r = 0; r < reps; ++r) {
      if (casync_forward_profiled(plan, xs[r % nsets], as[r % nsets], out, ws, B, flags, st, recs.data(), 1024, &nr)) {
        fprintf(stderr, "profiled: %s\n", casync_last_error());
        return 1;
      }
      if (acc.empty()) acc.assign(nr, 0.0);
      for (int i = 0; i < nr; ++i) acc[i] += recs[i].ms / reps;
    }
    double tot = 0;
    for (int i = 0; i < nr; ++i) tot += acc[i];
    printf("event-serialised total %.4f ms\n%-30s %9s %6s %8s %8s\n", tot, "launch", "us", "share", "GB/s", "TFLOP/s");
    for (int i = 0; i < nr; ++i)
      printf("%-30s %9.2f %5.1f%% %8.0f %8.1f\n", recs[i].name, acc[i] * 1e3, 100 * acc[i] / tot,
             recs[i].bytes / (acc[i] * 1e-3) / 1e9, recs[i].flops / (acc[i] * 1e-3) / 1e12);
  }
  casync_plan_destroy(plan);
  return 0;
}
