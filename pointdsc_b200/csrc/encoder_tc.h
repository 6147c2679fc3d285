// Tensor-core (tcgen05) encoder path — host-visible interface.  See encoder_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pdsc {

// Host pointers to one layer's folded fp32 weights (row-major [Cout][Cin]) and biases.
struct TcLayerHost {
  const float *w1, *b1, *wq, *bq, *wk, *bk, *wv, *bv, *wm0, *bm0, *wm1, *bm1, *wm2, *bm2;
};

// Device-resident operand images built once per pdsc_commit_params().
struct TcWeights {
  void* arena = nullptr;   // all images + biases
  size_t arena_bytes = 0;
  int num_layers = 0;
};

struct TcForwardArgs {
  int B, N, NS, in_dim, num_layers;
  int split;                 // 1: hi/lo operand split (3 products)   0: single 16-bit operands
  int fmt;                   // operand format of the kind::f16 MMAs: 0 = fp16, 1 = bf16
  const float* corr_pos;     // [B*N][in_dim]
  const float *l0w, *l0b;    // layer0 weights (fp32, device)
  const float* sc;           // [B][N][NS]
  float* feat;               // [B*N][128]  layer output / final features
  float* feat1;              // [B*N][128]  PointCN output (residual source)
  float* msg;                // [B*N][128]  attention output
  void* scratch;             // tc_scratch_bytes(B, N)
  int layer_tap;             // -1 or layer index to copy out
  float* layer_tap_out;
  int debug_layer;           // layer whose internals are decoded into debug_out
  float* debug_out;          // [5][B*N][128]: feat1, q (scaled by log2e/sqrt(C)), k, v, msg — or nullptr
  long long* timeline;       // nullptr or [2][16][4][8] clock64 stamps: chain<PCQ> and attention of layer `debug_layer`
  cudaEvent_t* attn_events;  // nullptr or 2 events per layer, recorded around the attention launch
};

int tc_build_weights(const TcLayerHost* layers, int num_layers, TcWeights* out);  // returns cudaError_t
void tc_free_weights(TcWeights* w);
size_t tc_scratch_bytes(int B, int N);
int tc_launches(int num_layers, int B, int N);
int tc_encoder_forward(const TcWeights& w, const TcForwardArgs& a, cudaStream_t st);  // returns cudaError_t

}  // namespace pdsc
