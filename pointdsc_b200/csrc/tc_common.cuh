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
  long long* dbg;        // optional timeline: clock64() stamps of CTA 0 (tools/tc_timeline.py)
};

// Two synchronisation rules of the tensor-core kernels (round 2; evidence: profiles/r02_determinism_campaign.txt).
//  1. COUNTED mbarriers and warps that run ahead.  A barrier that several independent warps arrive on must not be
//     reachable twice by one warp before its phase has completed: the second arrival is counted towards the CURRENT phase,
//     which then completes without the slowest warp.  This was the round-1 reproducibility bug: chain<PCQ> issues the
//     PointCN GEMM one tile ahead, so a fast epilogue warp could deliver tile t + 1's operand while a slow one was still
//     writing tile t's, on ONE barrier; the Q GEMM then read 32 rows (one lane quarter) of tensor memory before they were
//     written - about once in four forwards at B = 256.  Fix: one barrier per tile parity (tc_chain.cuh).  Every counted
//     barrier of these kernels is annotated with the reason a second arrival cannot overtake its phase.
//  2. Tensor-memory write-after-read.  PTX orders two tcgen05.mma only when they share accumulator and shape, so an MMA
//     that overwrites columns which an earlier MMA reads as its A operand waits for that MMA's COMPLETION through a commit
//     barrier (tcgen05.commit tracks all MMAs the thread issued before it); issue order alone is not relied upon.

// timeline stamp: slot = role * 64 + event (CTA 0 only, first 16 tiles)
#define PDSC_STAMP(dbg, it, role, ev)                                                       \
  do {                                                                                       \
    if ((dbg) && blockIdx.x == 0 && (it) < 16) (dbg)[((it) * 4 + (role)) * 8 + (ev)] = clock64(); \
  } while (0)

// same, for a caller that has already folded the "timeline on, CTA 0, my thread" test into one predicate
#define PDSC_STAMP1(dbg, it, role, ev)                                              \
  do {                                                                              \
    if ((it) < 16) (dbg)[((it) * 4 + (role)) * 8 + (ev)] = clock64();               \
  } while (0)

// issue one GEMM step: D[128 x NOUT] (+)= A[128 x 64*KP] * W[NOUT x 64*KP]^T, optionally as three hi/lo products.
// Fully unrolled; a descriptor differs from its neighbour only in the 14-bit start-address field, so each MMA costs
// one shift/mask per operand on the issuing thread.
template <int KP, int NOUT>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_panel_bytes,
                                           uint32_t b_hi, uint32_t b_lo, uint32_t b_panel_bytes, int split,
                                           uint32_t accumulate, int fmt) {
  const uint32_t idesc = idesc_f16kind(128, NOUT, fmt);
  constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
  uint32_t acc = accumulate;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t > 0 && !split) break;
    const uint32_t a = (t == 2) ? a_lo : a_hi;
    const uint32_t b = (t == 1) ? b_lo : b_hi;
#pragma unroll
    for (int p = 0; p < KP; ++p) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t alo = (((a + p * a_panel_bytes + ks * 32) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t blo = (((b + p * b_panel_bytes + ks * 32) >> 4) & 0x3FFFu) | (1u << 16);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(idesc), "r"(acc), "r"(kDescHi)
            : "memory");
        acc = 1;
      }
    }
  }
}

// 256-bit store (STG.256, sm_100+): a thread that owns a row writes whole 32-byte sectors of it, so a row-per-thread epilogue can
// store straight from registers without a shared-memory transposition
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                             uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(__cvta_generic_to_global(p)), "r"(a0), "r"(a1), "r"(a2),
               "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

}  // namespace pdsc
