// a4 + a5 — feature normalisation and the confidence head, one warp per correspondence.
//
// Reference: models/PointDSC.py:156  normed = F.normalize(features, p=2, dim=-1)   (eps 1e-12)
//            models/PointDSC.py:107-113, :171  confidence = Conv(128->32) ReLU Conv(32->32) ReLU Conv(32->1)
//            applied to the UN-normalised features.
// HBM-bound: reads 512 B and writes 516 B per correspondence; the 5.2 kMAC/point MLP rides along.
#include "common.cuh"
#include "kernels.h"

namespace pdsc {

// Lane = one correspondence (row), 32 rows per warp pass.  The warp's 32 rows arrive with cp.async (16 bytes per lane, one
// full 512-byte row per instruction) into a ROW-major shared-memory tile with a row stride of 132 floats: lane i then reads
// its own row four channels at a time (LDS.128; 16-byte chunk index lane * 33 + c / 4 is distinct modulo 8 within every
// quarter warp: conflict free), runs the whole MLP of that row with the 32 hidden accumulators in registers while the
// weights arrive as warp-wide BROADCAST 16-byte loads (32 FMAs per 8 weight loads), and the normalised rows leave the
// tile again as full 512-byte rows.  Eight warps per CTA, one tile each: while a warp waits for its next tile the other
// warp of its scheduler computes.  Accumulation order is the reference's: bias first, then ascending input channel, one
// fp32 FMA each — issued two outputs at a time as FFMA2 (the activation is the instruction's scalar operand).  (Round 1 staged the tile transposed through 4-way conflicting scalar stores with 2 x 4 warps per SM and
// no overlap of load and compute: 0.20 ms for 262 MB.)
constexpr int kHeadWarps = 8;
constexpr int kHeadStride = 132;       // floats per staged row

__device__ __forceinline__ void head_cp_async_16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kHeadWarps * 32) head_kernel(const float* __restrict__ feat, HeadWeights w,
                                                                float* __restrict__ normed, float* __restrict__ conf,
                                                                long long rows, int want_conf) {
  extern __shared__ __align__(16) float hsm[];
  float* w0t = hsm;                    // [c][o]  128 x 32
  float* w2t = w0t + kC * 32;          // [c][o]   32 x 32
  float* b0s = w2t + 32 * 32;
  float* b2s = b0s + 32;
  float* w4s = b2s + 32;
  float* tiles = w4s + 32;             // [warp][32 * kHeadStride]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (want_conf) {
    for (int i = tid; i < kC * 32; i += kHeadWarps * 32) w0t[i] = w.w0t[i];
    for (int i = tid; i < 32 * 32; i += kHeadWarps * 32) w2t[i] = w.w2t[i];
    if (tid < 32) { b0s[tid] = w.b0[tid]; b2s[tid] = w.b2[tid]; w4s[tid] = w.w4[tid]; }
  }
  __syncthreads();
  const float b4 = want_conf ? w.b4[0] : 0.f;
  float* T = tiles + (size_t)warp * 32 * kHeadStride;
  const uint32_t t_base = (uint32_t)__cvta_generic_to_shared(T);
  const long long ntiles = (rows + 31) / 32;
  for (long long t = (long long)blockIdx.x * kHeadWarps + warp; t < ntiles; t += (long long)gridDim.x * kHeadWarps) {
    const long long r0 = t * 32;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      if (r0 + i < rows) head_cp_async_16(t_base + (uint32_t)((i * kHeadStride + lane * 4) * 4), feat + (r0 + i) * kC + lane * 4);
      else *reinterpret_cast<float4*>(T + i * kHeadStride + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    // this lane's row: squared norm (+ hidden layer 1 when the confidence is wanted)
    const float* my = T + lane * kHeadStride;
    float ss = 0.f;
    float h1[32];
    if (want_conf) {
#pragma unroll
      for (int o = 0; o < 32; ++o) h1[o] = b0s[o];
#pragma unroll 1
      for (int c4 = 0; c4 < kC; c4 += 4) {
        const float4 fv = *reinterpret_cast<const float4*>(my + c4);
        const float fr[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float f = fr[q];
          ss = fmaf(f, f, ss);
#pragma unroll
          for (int o4 = 0; o4 < 8; ++o4) {
            const float4 wv = *reinterpret_cast<const float4*>(w0t + (c4 + q) * 32 + o4 * 4);
            ffma2(h1[o4 * 4 + 0], h1[o4 * 4 + 1], f, wv.x, wv.y);      // two outputs per instruction, each a plain fp32 FMA
            ffma2(h1[o4 * 4 + 2], h1[o4 * 4 + 3], f, wv.z, wv.w);
          }
        }
      }
    } else {
#pragma unroll 4
      for (int c4 = 0; c4 < kC; c4 += 4) {
        const float4 fv = *reinterpret_cast<const float4*>(my + c4);
        ss = fmaf(fv.x, fv.x, ss); ss = fmaf(fv.y, fv.y, ss); ss = fmaf(fv.z, fv.z, ss); ss = fmaf(fv.w, fv.w, ss);
      }
    }
    const float den = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize: x / max(||x||, eps)
    if (want_conf) {
      float h2[32];
#pragma unroll
      for (int o = 0; o < 32; ++o) h2[o] = b2s[o];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float a = fmaxf(h1[c], 0.f);
#pragma unroll
        for (int o4 = 0; o4 < 8; ++o4) {
          const float4 wv = *reinterpret_cast<const float4*>(w2t + c * 32 + o4 * 4);
          ffma2(h2[o4 * 4 + 0], h2[o4 * 4 + 1], a, wv.x, wv.y);
          ffma2(h2[o4 * 4 + 2], h2[o4 * 4 + 3], a, wv.z, wv.w);
        }
      }
      float o = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) o = fmaf(fmaxf(h2[c], 0.f), w4s[c], o);
      if (r0 + lane < rows) conf[r0 + lane] = o + b4;
    }
    // normalised rows out, one full 512-byte row per store instruction
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const float d = __shfl_sync(0xffffffffu, den, i);
      if (r0 + i < rows) {
        const float4 fv = *reinterpret_cast<const float4*>(T + i * kHeadStride + lane * 4);
        *reinterpret_cast<float4*>(normed + (r0 + i) * kC + lane * 4) = make_float4(fv.x / d, fv.y / d, fv.z / d, fv.w / d);
      }
    }
    __syncwarp();
  }
}

void launch_head(const float* feat, const HeadWeights& w, float* normed, float* conf, long long rows, int want_conf,
                 cudaStream_t st) {
  long long blocks = ((rows + 31) / 32 + kHeadWarps - 1) / kHeadWarps;
  const long long max_blocks = device_sm_count();
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  constexpr int kSmem = (kC * 32 + 32 * 32 + 96 + kHeadWarps * 32 * kHeadStride) * (int)sizeof(float);
  ensure_dynamic_smem(reinterpret_cast<const void*>(head_kernel), kSmem);
  head_kernel<<<(unsigned)blocks, kHeadWarps * 32, kSmem, st>>>(feat, w, normed, conf, rows, want_conf);
}

}  // namespace pdsc
