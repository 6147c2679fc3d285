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

// Two independent fp32 FMAs in one instruction (Blackwell FFMA2, PTX fma.rn.f32x2): d0 += a * b0, d1 += a * b1, each rounded
// exactly like a scalar fmaf.  ptxas folds the broadcast of `a` into the instruction's scalar-operand form and needs no moves
// when (d0, d1) and (b0, b1) are neighbouring registers (array elements 2p, 2p + 1; the halves of a float4).
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a), "f"(b0), "f"(b1));
}

// element-wise pairs (FADD2 / FMUL2 / FFMA2): every lane of the pair is rounded exactly like the scalar operation
__device__ __forceinline__ float2 fsub2_scalar(float a, float2 b) {      // (a - b.x, a - b.y)
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fsub2_pair_scalar(float2 a, float b) {   // (a.x - b, a.y - b)
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 ffma2_pair(float2 a, float2 b, float2 c) {   // a * b + c
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
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
// the same for two values at once: every lane runs the identical rounded sequence (FMUL2, FMUL2, FFMA2, FFMA2, FADD2, max)
__device__ __forceinline__ float2 div_by_const2(float2 x, float c, float rc) {
  float2 q0 = fmul2(x, make_float2(rc, rc));            // q0 = RN(x rc)
  float2 r = x;
  ffma2(r.x, r.y, -c, q0.x, q0.y);                      // r = x - q0 c, exact
  ffma2(q0.x, q0.y, rc, r.x, r.y);                      // q = RN(q0 + r rc)
  return q0;
}
__device__ __forceinline__ float2 consistency_rc2(float2 d, float s2, float rc_s2) {
  const float2 one_minus = fsub2_scalar(1.0f, div_by_const2(fmul2(d, d), s2, rc_s2));
  return make_float2(fmaxf(one_minus.x, 0.0f), fmaxf(one_minus.y, 0.0f));
}
// length3_pow() of two difference vectors side by side
__device__ __forceinline__ float2 length3_pow2(float2 dx, float2 dy, float2 dz) {
  const float2 s = fadd2(fadd2(fmul2(dx, dx), fmul2(dy, dy)), fmul2(dz, dz));
  return make_float2(__fsqrt_rn(s.x), __fsqrt_rn(s.y));
}

}  // namespace pdsc
