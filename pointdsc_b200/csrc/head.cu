// a4 + a5 — feature normalisation and the confidence head, one warp per correspondence.
//
// Reference: models/PointDSC.py:156  normed = F.normalize(features, p=2, dim=-1)   (eps 1e-12)
//            models/PointDSC.py:107-113, :171  confidence = Conv(128->32) ReLU Conv(32->32) ReLU Conv(32->1)
//            applied to the UN-normalised features.
// HBM-bound: reads 512 B and writes 516 B per correspondence; the 5.2 kMAC/point MLP rides along.
#include "common.cuh"
#include "kernels.h"

namespace pdsc {

__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ feat, HeadWeights w,
                                                   float* __restrict__ normed, float* __restrict__ conf,
                                                   long long rows, int want_conf) {
  __shared__ float w0t[kC * 32];   // [c][o]
  __shared__ float w2t[32 * 32];   // [c][o]
  __shared__ float frow[8][kC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (want_conf) {
    for (int i = tid; i < kC * 32; i += 256) w0t[i] = w.w0t[i];
    for (int i = tid; i < 32 * 32; i += 256) w2t[i] = w.w2t[i];
  }
  __syncthreads();
  const float b0 = want_conf ? w.b0[lane] : 0.f, b2 = want_conf ? w.b2[lane] : 0.f;
  const float w4 = want_conf ? w.w4[lane] : 0.f, b4 = want_conf ? w.b4[0] : 0.f;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const float4 f = *reinterpret_cast<const float4*>(feat + r * kC + lane * 4);
    const float ss = warp_sum(f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w);
    const float den = fmaxf(sqrtf(ss), 1e-12f);
    *reinterpret_cast<float4*>(normed + r * kC + lane * 4) = make_float4(f.x / den, f.y / den, f.z / den, f.w / den);
    if (!want_conf) continue;
    __syncwarp();
    *reinterpret_cast<float4*>(&frow[warp][lane * 4]) = f;
    __syncwarp();
    float h1 = b0;
#pragma unroll 8
    for (int c = 0; c < kC; ++c) h1 = fmaf(frow[warp][c], w0t[c * 32 + lane], h1);
    h1 = fmaxf(h1, 0.f);
    float h2 = b2;
#pragma unroll
    for (int c = 0; c < 32; ++c) h2 = fmaf(__shfl_sync(0xffffffffu, h1, c), w2t[c * 32 + lane], h2);
    h2 = fmaxf(h2, 0.f);
    const float o = warp_sum(h2 * w4);
    if (lane == 0) conf[r] = o + b4;
  }
}

void launch_head(const float* feat, const HeadWeights& w, float* normed, float* conf, long long rows, int want_conf,
                 cudaStream_t st) {
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  head_kernel<<<(unsigned)blocks, 256, 0, st>>>(feat, w, normed, conf, rows, want_conf);
}

}  // namespace pdsc
