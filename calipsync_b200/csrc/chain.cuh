// Layer-program ("chain") kernel: ONE persistent launch runs a sequence of low-resolution layers -- tcgen05 GEMMs
// (1x1 convs, Linear, implicit-GEMM 3x3, upsample+concat producer) and the depthwise 3x3 convs between them --
// as a list of work items (layer, row tile, column tile).  Kernel boundaries are replaced by per-row-tile
// completion counters in global memory: an item starts as soon as the row tiles of the producing layer that it
// reads are complete, so the tail of one layer overlaps the head of the next, the epilogue of one tile overlaps the
// main loop of the CTA's next item whatever layer that belongs to, and there is no per-layer launch, prologue or
// wave quantisation.  See chain.cu.
#pragma once
#include <vector>

#include "gemm_tc.cuh"

namespace casync {

struct ChainDep {
  int mode;    // 0: none.  1: same row space as the producer, +- `a` halo rows.  2: through frames: this layer has `a`
  int layer;   //    rows per frame, the producer `b`; a row tile needs every producer row of the frames it touches
  int a, b;
};

enum ChainKind : int { CK_GEMM = 0, CK_DW = 1 };

// Device-visible layer descriptor (array in global memory, written by the host before the launch).
struct alignas(128) ChainLayer {
  unsigned char tm[128];   // CUtensorMap: GEMM with a plain row-major A operand: box 64 x 128 rows, SWIZZLE_128B;
                           // depthwise stride 1: the hidden tensor, box 64 x (130 + 2W) rows, unswizzled
  GemmArgs g;              // GEMM: as launch_gemm.  Depthwise: A = hidden NHWC [B,Hin,Win,K], C = out [M,K], K = channels,
                           // M = output rows, W = taps + bias as bf16 [K/8][10][8], stride
  int kind;                // ChainKind
  int BN, NT, MT;          // GEMM: column tile / column tiles / row tiles.  Depthwise: NT = groups of <= 256 channels
  int item0, items;        // work items [item0, item0 + items): item = mt * NT + nt
  int cnt0;                // first completion counter (one per row tile)
  int need;                // arrivals that complete a row tile (= NT)
  int out_rows;            // rows of the output (= g.M)
  int res_late;            // the residual operand is written inside this program: it may only be read once the
                           // accumulator is complete (the A-side dependency wait then covers it transitively)
  int pad_[2];
  ChainDep dep[2];
};

// Host-side builder: collects layers, derives the dependencies from the buffers they read and write, and launches.
class Chain {
 public:
  void reset();
  bool empty() const { return layers_.empty(); }
  size_t size() const { return layers_.size(); }
  // returns 0, or -1 when the layer cannot join the program (caller flushes and retries, or launches it alone)
  int add_gemm(const GemmArgs& a);
  // wdp: depthwise taps + folded-BN bias as bf16, [C/8][10][8] (9 taps, then the bias, per 8-channel chunk)
  int add_dw(const __nv_bfloat16* in, __nv_bfloat16* out, const uint8_t* wdp, int batch, int H, int W, int C, int stride);
  // bytes of device scratch (descriptors + counters) the current program needs
  size_t scratch_bytes() const;
  // uploads the descriptors to `scratch` (256-byte aligned device memory; skipped when `last`, the host copy of what
  // the device holds, is identical), zeroes the counters behind them, launches, and clears the program
  int launch(void* scratch, size_t scratch_cap, std::vector<unsigned char>& last, cudaStream_t st);
  const GemmArgs& gemm(int i) const { return layers_[i].g; }
  int kind(int i) const { return layers_[i].kind; }

 private:
  struct Out {
    const unsigned char* base;
    size_t bytes;
    int ld, ncols, layer;
  };
  int find_deps(ChainLayer& L, const void* ptr, int ld, int ncols, size_t bytes, int mode, int a, int b);
  std::vector<ChainLayer> layers_;
  std::vector<Out> outs_;
  int items_ = 0, counters_ = 0;
};

int chain_init();   // set the kernel's shared-memory attribute (once per process)
void chain_dbg_report();   // developer timing (CASYNC_CHAIN_DBG=1): per-role cycle counters to stderr

constexpr int kChainMaxLayers = 44;

}  // namespace casync
