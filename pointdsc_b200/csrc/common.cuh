// Shared device helpers for the pointdsc_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdsc {

constexpr int kC = 128;          // feature channels (num_channels of the released models)
constexpr int kMaxK = 128;       // largest supported NSM neighbourhood size
constexpr int kMaxIters = 16;    // largest supported power-iteration cap

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Euclidean length exactly as the reference's `torch.norm(x, dim=-1)` produces it on CPU (fp32 accumulate,
// compiled with FMA contraction: x*x, then fma(y,y,.), then fma(z,z,.), IEEE sqrt) — verified bit for bit
// against torch 2.11 on 4M vectors.  Used for src_dist / SC (reference models/PointDSC.py:151-152).
__device__ __forceinline__ float length3(float dx, float dy, float dz) {
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// Euclidean length as `((d) ** 2).sum(-1) ** 0.5` produces it (three rounded squares, (x+y)+z, sqrt):
// the form used inside cal_seed_trans (reference models/PointDSC.py:268).
__device__ __forceinline__ float length3_pow(float dx, float dy, float dz) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// max(0, 1 - d^2 / s2) with every operation rounded separately (PointDSC.py:153, :270).
__device__ __forceinline__ float consistency(float d, float s2) {
  return fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(d, d), s2)), 0.0f);
}

// x / c for a constant c whose correctly rounded reciprocal rc = RN(1 / c) is known: q0 = RN(x rc), the EXACT residual
// r = x - q0 c (one FMA), q = RN(q0 + r rc).  This is the refinement step div.rn itself ends with (Markstein): the result is
// the correctly rounded quotient for the operand ranges of this engine (no overflow / underflow: x <= 1e4, c in [1e-3, 1e2]),
// in 3 instructions instead of the ~12 (reciprocal approximation, two refinements, range check, slow-path call) of a general
// IEEE division.  The issue-bound SC kernel and the NSM compatibility block divide hundreds of millions of times by sigma^2.
__device__ __forceinline__ float div_by_const(float x, float c, float rc) {
  const float q0 = __fmul_rn(x, rc);
  const float r = __fmaf_rn(-q0, c, x);
  return __fmaf_rn(r, rc, q0);
}
__device__ __forceinline__ float consistency_rc(float d, float s2, float rc_s2) {
  return fmaxf(__fsub_rn(1.0f, div_by_const(__fmul_rn(d, d), s2, rc_s2)), 0.0f);
}

}  // namespace pdsc
