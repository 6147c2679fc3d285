// Shared definitions of the tensor-core encoder kernels (included by encoder_tc.cu only).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdsc {
using namespace ptx;

// ---- weight arena layout (bytes, per layer) ------------------------------------------------------------
constexpr size_t kW1 = 0, kWq = 65536, kWk = 131072, kWv = 196608, kWm0 = 262144, kWm1 = 294912, kWm2 = 311296,
                 kBias = 344064, kLayerBytes = 348160;
// bias block (floats): b1[128] bq[128] bk[128] bv[128] bm0[64] bm1[64] bm2[128]
constexpr int kB1 = 0, kBq = 128, kBk = 256, kBv = 384, kBm0 = 512, kBm1 = 576, kBm2 = 640, kBiasFloats = 768;

constexpr float kQScale = 1.4426950408889634f / 11.313708498984761f;  // log2(e) / sqrt(128)

enum ChainMode { kPCQ = 0, kKV = 1, kMSG = 2 };

struct ChainArgs {
  long long rows;        // B * N
  int N, QT, KT, split;
  const float* in;       // [rows][128] fp32 A operand
  const float* res;      // MSG: feat1 (residual)
  float* out_f32;        // PCQ: feat1, MSG: feat
  uint8_t* qimg;
  uint8_t* kvimg;
  const uint8_t* wimg;   // this kernel's weight images (contiguous)
  const float* bias;     // the layer's bias block
  int wbytes;            // bytes of weight images to stage
};

// issue one GEMM step: D[128 x Nout] (+)= A[128 x K] * W[Nout x K]^T, optionally as three hi/lo products
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_panel_bytes,
                                           uint32_t b_hi, uint32_t b_lo, uint32_t b_panel_bytes, int K, int Nout,
                                           int split, uint32_t accumulate, int fmt) {
  const uint32_t idesc = idesc_f16kind(128, Nout, fmt);
  const int terms = split ? 3 : 1;
  uint32_t acc = accumulate;
  for (int t = 0; t < terms; ++t) {
    const uint32_t a = (t == 2) ? a_lo : a_hi;
    const uint32_t b = (t == 1) ? b_lo : b_hi;
    for (int p = 0; p < K / 64; ++p) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        mma_bf16(d_tmem, smem_desc_sw128(a + p * a_panel_bytes + ks * 32), smem_desc_sw128(b + p * b_panel_bytes + ks * 32),
                 idesc, acc);
        acc = 1;
      }
    }
  }
}

// 8 consecutive fp32 values -> one 16-byte hi chunk and one 16-byte lo chunk
template <int FMT>
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  split_pair<FMT>(x[0], x[1], hi.x, lo.x);
  split_pair<FMT>(x[2], x[3], hi.y, lo.y);
  split_pair<FMT>(x[4], x[5], hi.z, lo.z);
  split_pair<FMT>(x[6], x[7], hi.w, lo.w);
}

}  // namespace pdsc
