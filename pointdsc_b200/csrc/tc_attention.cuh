// tc_attention: SC-weighted flash attention on tcgen05 (included by encoder_tc.cu only).
//
// Reference: models/PointDSC.py:39-42
//     P = softmax_j( SC_ij * (q_i . k_j) / sqrt(C) ),   msg_i = sum_j P_ij v_j          (heads = 1, C = 128)
// SC multiplies the logit (it is not a mask): SC_ij = 0 leaves logit 0, which still takes softmax mass.
//
// This header holds what the attention kernel (tc_attention_p.cuh, persistent CTAs) is built from: the argument block,
// the SC tile loader, the exponential, and the two MMA issue helpers (S = Q K^T with Q in tensor memory, O += P V with P in
// tensor memory and V read as an MN-major B operand).
#pragma once
#include "tc_common.cuh"

namespace pdsc {

struct AttnArgs {
  int N, NS, QT, KT, split;
  const uint8_t* qimg;
  const uint8_t* kvimg;
  const float* sc;    // tiled: [B][KT][QT][16 key groups][128 queries][4 keys]
  float* msg;
  long long* dbg;
  int items;          // work items of the launch: B * QT * splits
  // key split (small calls only, see encoder_tc.cu): a work item is (set, query tile, split) and covers key tiles
  // [split * TS, split * TS + TS); tiles beyond KT are "virtual" (fully masked).  splits == 1: TS == KT, one item per query tile.
  int splits, TS;
  float* part_o;      // [items][128][128] unnormalised O of every work item           (splits > 1)
  float* part_ml;     // [items][128][2]   its final reference maximum (log2 units) and row sum
};

constexpr int kAttnThreads = 320;
constexpr float kRescaleThreshold = 8.0f;              // log2 units: P < 2^8 before the reference max advances

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_all_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// streaming read-only load: every SC element is used once per CTA, keep it out of L1
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// this thread's (query row r) 64 SC values of one tile [16 key groups][128 queries][4 keys]: 16 coalesced 16-byte loads
__device__ __forceinline__ void load_sc_tile(float (&sc)[64], const float* tile, int r) {
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const float4 v = ldg_stream4(tile + (g * 128 + r) * 4);
    sc[4 * g] = v.x; sc[4 * g + 1] = v.y; sc[4 * g + 2] = v.z; sc[4 * g + 3] = v.w;
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// D[128 x NOUT] (+)= A[128 x 64 KP] * B[NOUT x 64 KP]^T with A in TENSOR MEMORY (lane = row, 32-bit column c = K elements
// 2c | 2c+1; hi image at a_hi, lo image at a_lo) and B K-major SWIZZLE_128B panels (64 K elements each) in shared memory.
// With A in tensor memory the tensor core only streams B from shared memory: an N = 64 step costs 32 cycles instead of
// the ~64 it costs when the 4 KB A slice is re-read from shared memory as well.
template <int KP, int NOUT>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t b_panel_bytes, int split, uint32_t accumulate, int fmt) {
  const uint32_t idesc = idesc_f16kind(128, NOUT, fmt);
  constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
  uint32_t acc = accumulate;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t > 0 && !split) break;
    const uint32_t at = (t == 2) ? a_lo : a_hi;
    const uint32_t b = (t == 1) ? b_lo : b_hi;
#pragma unroll
    for (int p = 0; p < KP; ++p) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t blo = (((b + p * b_panel_bytes + ks * 32) >> 4) & 0x3FFFu) | (1u << 16);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(at + (p * 4 + ks) * 8), "r"(blo), "r"(idesc), "r"(acc), "r"(kDescHi)
            : "memory");
        acc = 1;
      }
    }
  }
}

// O[128 x 128] (+)= P[128 x 64] * V[64 x 128] with P in tensor memory and V (rows = keys, the K image format: two 64-channel
// panels of 64 rows x 128 B) read as an MN-MAJOR B operand: idesc bit 16, LBO = 8192 B between the two 64-channel atoms,
// SBO = 1024 B between 8-key groups, 2048 B per K = 16 step (pinned by tools/mn_major_probe.cu).
__device__ __forceinline__ void issue_pv_mn(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t v_hi, uint32_t v_lo, int split,
                                            uint32_t accumulate, int fmt) {
  const uint32_t idesc = idesc_f16kind(128, 128, fmt) | (1u << 16);
  constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
  uint32_t acc = accumulate;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t > 0 && !split) break;
    const uint32_t at = (t == 2) ? a_lo : a_hi;
    const uint32_t b = (t == 1) ? v_lo : v_hi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t blo = (((b + ks * 2048) >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
          "mov.b64 db, {%2, %5};\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
          ::"r"(d_tmem), "r"(at + ks * 8), "r"(blo), "r"(idesc), "r"(acc), "r"(kDescHi)
          : "memory");
      acc = 1;
    }
  }
}

}  // namespace pdsc
