// tc_attention: SC-weighted flash attention on tcgen05 (included by encoder_tc.cu only).
//
// Reference: models/PointDSC.py:39-42
//     P = softmax_j( SC_ij * (q_i . k_j) / sqrt(C) ),   msg_i = sum_j P_ij v_j          (heads = 1, C = 128)
// SC multiplies the logit (it is not a mask): SC_ij = 0 leaves logit 0, which still takes softmax mass.
//
// One CTA = 128 queries of one set, looping over 64-key tiles.  Warp roles (320 threads):
//   warp 0      loader : bulk async copies (TMA engine) of the ready-made K / V operand images (K by pairs of tiles, ring of 2; V ring of 2)
//   warp 1      MMA    : S_j = Q K_j^T into one of four TMEM buffers, issued up to three tiles ahead;
//                        O += P_j V_j with P_j read FROM TENSOR MEMORY (A operand in TMEM); owns the TMEM allocation
//   warps 2-5   softmax group 0: even key tiles          warps 6-9  softmax group 1: odd key tiles
//               thread = one query row (TMEM lane) and all 64 logits of the tile, so a tile needs no cross-thread
//               reduction; the two groups work on consecutive tiles half a period apart, so one group's exponentials
//               (MUFU) run under the other group's conversions (ALU).  Logits (log2 domain; Q carries log2e/sqrt(C))
//               = S * SC with SC read from the tiled layout of sc_matrix.cu at compile-time offsets.  The running
//               reference maximum of a row is shared by the two threads that own it (one per group) through shared
//               memory, tile by tile; it only advances when the row maximum grew by > 8 (FA4-style lazy rescale), so
//               O in TMEM is rescaled rarely.  P = ex2(l - ref) is split hi/lo (16-bit) and written over its own S
//               tile in TMEM: it never touches shared memory.
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
  int items;          // B * QT work items (persistent kernel)
};

constexpr int kAttnThreads = 320;
// K is staged per PAIR of 64-key tiles so that S = Q K^T runs as N = 128 MMAs (an N = 64 MMA is bound by the delivery of
// its 4 KB A slice, ~48 cycles instead of 32): stage = [hi p0 | hi p1 | lo p0 | lo p1], each panel 128 rows x 128 B with
// the first tile of the pair in rows 0-63 and the second in rows 64-127.  The Q image lands in K stage 1 first (it is
// moved to tensor memory before the second pair is needed); the output tile is staged in K stage 0 at the end.
constexpr int kAttnKStages = 2, kAttnVStages = 2;
constexpr int kAttnK = 0, kAttnQ = 65536, kAttnV = kAttnKStages * 65536, kAttnBars = kAttnV + kAttnVStages * 32768;
constexpr int kAttnOut = 0;
constexpr int kAttnRef = kAttnBars + 256;              // float ref[2][128], lsum[2][128]
constexpr int kAttnSmemTc = kAttnRef + 2048;           // 198,912 B
static_assert(kAttnBars == 196608, "smem map");
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

template <int FMT>
__global__ void __launch_bounds__(kAttnThreads, 1) tc_attention_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttnBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);   // bars[24] holds the TMEM base, bars[25] is q_tmem
  float* ref_s = reinterpret_cast<float*>(smem + kAttnRef);   // [2][128] reference maximum after tile j (slot j & 1)
  float* lsum_s = ref_s + 256;                                // [2][128] per-group row sums (epilogue)
  const uint32_t s0 = smem_u32(smem);
  const uint32_t q_full = smem_u32(bars + 0);
  const uint32_t k_full = smem_u32(bars + 1), k_empty = smem_u32(bars + 4);     // [2] (per pair of key tiles)
  const uint32_t v_full = smem_u32(bars + 7), v_empty = smem_u32(bars + 9);     // [2]
  const uint32_t s_full = smem_u32(bars + 11), p_full = smem_u32(bars + 15);    // s_full [2] per pair, p_full [4] per tile
  const uint32_t pv_done = smem_u32(bars + 19), ref_ready = smem_u32(bars + 21);  // [2]
  const uint32_t o_done = smem_u32(bars + 23), q_tmem = smem_u32(bars + 25);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.QT, qt = blockIdx.x % a.QT;
  const int T = a.KT;
  const bool life = a.dbg != nullptr && (blockIdx.x == 0 || blockIdx.x == 1000) && tid == 64;   // row 0: block 0, row 1: block 1000   // CTA lifetime stamps: role 3 of tile row 0 / 1
  if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 0);

  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(q_full, 1);
    mbar_init(o_done, 1);
    mbar_init(q_tmem, 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1);
      mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
      mbar_init(pv_done + 8 * i, 1); mbar_init(ref_ready + 8 * i, 128);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(s_full + 8 * i, 1); mbar_init(p_full + 8 * i, 128); }   // s_full: [0..1] used
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 1);
  const uint32_t tO = tmem + 256;   // S/P buffer i at +64 i (i = tile & 3), O at +256 (128 fp32 columns)
  const uint32_t tQ = tmem + 384;   // Q: hi image (64 columns = 128 channels) at +384, lo image at +448

  if (warp == 0) {
    // ===================================== loader =====================================
    if (lane == 0) {
      const uint8_t* qsrc = a.qimg + ((size_t)b * a.QT + qt) * 65536;
      mbar_expect_tx(q_full, a.split ? 65536u : 32768u);
      bulk_g2s(s0 + kAttnQ, qsrc, 32768u, q_full);
      if (a.split) bulk_g2s(s0 + kAttnQ + 32768, qsrc + 32768, 32768u, q_full);
      const uint8_t* kv = a.kvimg + (size_t)b * a.KT * 65536;
      const int TP = (T + 1) >> 1;
      // two independent streams (K by pairs, two pairs deep; V by tiles, two deep): a full V ring never holds back K
      int kp = 0, vj = 0;
      while (kp < TP || vj < T) {
        bool progress = false;
        if (kp < TP) {
          const int st = kp & 1, use = kp >> 1;
          // stage 1 holds the Q image until the softmax threads have moved it to tensor memory
          const bool free = (use == 0) ? (st == 0 || mbar_test(q_tmem, 0)) : mbar_test(k_empty + 8 * st, (uint32_t)((use - 1) & 1));
          if (free) {
            const int ntiles = (2 * kp + 1 < T) ? 2 : 1;
            mbar_expect_tx(k_full + 8 * st, (a.split ? 32768u : 16384u) * ntiles);
            for (int hh = 0; hh < ntiles; ++hh) {
              const uint8_t* src = kv + (size_t)(2 * kp + hh) * 65536;
              const uint32_t dst = s0 + kAttnK + st * 65536 + hh * 8192;
              bulk_g2s(dst, src, 8192u, k_full + 8 * st);                          // hi, channels 0-63
              bulk_g2s(dst + 16384, src + 8192, 8192u, k_full + 8 * st);           // hi, channels 64-127
              if (a.split) {
                bulk_g2s(dst + 32768, src + 16384, 8192u, k_full + 8 * st);        // lo
                bulk_g2s(dst + 49152, src + 24576, 8192u, k_full + 8 * st);
              }
            }
            ++kp;
            progress = true;
          }
        }
        if (vj < T) {
          const int st = vj & 1, use = vj >> 1;
          if (use == 0 || mbar_test(v_empty + 8 * st, (uint32_t)((use - 1) & 1))) {
            const uint32_t half = a.split ? 32768u : 16384u;
            mbar_expect_tx(v_full + 8 * st, half);
            bulk_g2s(s0 + kAttnV + st * 32768, kv + (size_t)vj * 65536 + 32768, half, v_full + 8 * st);
            ++vj;
            progress = true;
          }
        }
        if (!progress) __nanosleep(64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The whole warp runs the control flow (waits included); one elected lane issues the MMAs and commits.
    const bool leader = elect_one();
    const bool stamp_mma = leader && a.dbg != nullptr && blockIdx.x == 0;
    mbar_wait(q_tmem, 0);   // Q (hi | lo images, 64 columns each) is resident in tensor memory
    tc_fence_after();
    auto issue_qk_pair = [&](int p) {   // S tiles 2p, 2p+1 (columns 128 (p & 1) ...) = Q K^T over 128 keys
      const int st = p & 1, use = p >> 1;
      mbar_wait(k_full + 8 * st, (uint32_t)(use & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t kb = s0 + kAttnK + st * 65536;
        issue_gemm_ts<2, 128>(tmem + 128 * st, tQ, tQ + 64, kb, kb + 32768, 16384, a.split, 0, FMT);
        mma_commit(s_full + 8 * st);
        mma_commit(k_empty + 8 * st);
      }
    };
    const int TP = (T + 1) >> 1;
    for (int p = 0; p < TP && p < 2; ++p) issue_qk_pair(p);
    for (int j = 0; j < T; ++j) {
      const int vs = j & 1;
      if (stamp_mma) PDSC_STAMP1(a.dbg, j, 0, 0);
      mbar_wait(p_full + 8 * (j & 3), (uint32_t)((j >> 2) & 1));
      if (stamp_mma) PDSC_STAMP1(a.dbg, j, 0, 1);
      mbar_wait(v_full + 8 * vs, (uint32_t)((j >> 1) & 1));
      if (stamp_mma) PDSC_STAMP1(a.dbg, j, 0, 2);
      tc_fence_after();
      if (leader) {
        const uint32_t vb = s0 + kAttnV + vs * 32768;
        const uint32_t tP = tmem + 64 * (j & 3);   // P_j: hi image in columns [0,32), lo image in [32,64) of its S tile
        issue_pv_mn(tO, tP, tP + 32, vb, vb + 16384, a.split, j > 0 ? 1u : 0u, FMT);
        mma_commit(pv_done + 8 * vs);
        mma_commit(v_empty + 8 * vs);
      }
      if (stamp_mma) PDSC_STAMP1(a.dbg, j, 0, 3);
      // after PV of the second tile of pair p, pair p + 2 may overwrite that S/P buffer (in-order execution)
#if PDSC_STRICT_TMEM_WAR
      if ((j & 1) && (j >> 1) + 2 < TP) { mbar_wait(pv_done + 8 * vs, (uint32_t)((j >> 1) & 1)); tc_fence_after(); }
#endif
      if ((j & 1) && (j >> 1) + 2 < TP) issue_qk_pair((j >> 1) + 2);
    }
    if (leader) mma_commit(o_done);
    __syncwarp();
  } else {
    // ===================================== softmax =====================================
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int g = (warp - 2) >> 2;           // group: tiles j with (j & 1) == g
    const int r = q4 * 32 + lane;            // query row within the tile == TMEM lane
    const int gt = (warp - 2 - 4 * g) * 32 + lane;   // thread index within the group
    const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
    // SC tiles of this CTA: sc_t[b][j][qt][16 key groups][128 queries][4 keys]; thread r reads its 64 values of tile j at the
    // compile-time offsets g * 2048 B (16 float4 loads) from one per-tile pointer, coalesced over the 128 rows.
    const size_t tile_stride = (size_t)a.QT << 13;
    const float* sc_cta = a.sc + ((((size_t)b * a.KT) * a.QT + qt) << 13);
    const float* sc_line = sc_cta + gt * 32;   // two 128-byte lines of each 32 KB tile per thread (L2 prefetch)
    const bool ragged = (a.N & 63) != 0;
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && gt == 0;
    float my_ref = -INFINITY, l_sum = 0.f;

    float sc[64];
    if (g < T) {
      load_sc_tile(sc, sc_cta + (size_t)g * tile_stride, r);
    }
    if (g + 2 < T) { prefetch_l2(sc_line + (size_t)(g + 2) * tile_stride); prefetch_l2(sc_line + (size_t)(g + 2) * tile_stride + 4096); }
    // Q image: shared memory (landed by bulk copy) -> tensor memory, this thread's row; group 0 moves the hi image,
    // group 1 the lo image.  Logical 16-byte chunk c of row r sits at physical chunk c ^ (r & 7) of its 128-byte row.
    mbar_wait(q_full, 0);
    if (g == 0 || a.split) {
      const uint8_t* qrow = smem + kAttnQ + g * 32768 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int pnl = 0; pnl < 2; ++pnl) {
        uint32_t qv[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(qrow + pnl * 16384 + ((c ^ (r & 7)) << 4));
          qv[4 * c] = v.x; qv[4 * c + 1] = v.y; qv[4 * c + 2] = v.z; qv[4 * c + 3] = v.w;
        }
        tmem_st32(tQ + lane_base + 64 * g + 32 * pnl, qv);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(q_tmem);
    if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 2);
    for (int j = g; j < T; j += 2) {
      const uint32_t tS = tmem + 64 * (j & 3) + lane_base;
      if (j + 4 < T) { prefetch_l2(sc_line + (size_t)(j + 4) * tile_stride); prefetch_l2(sc_line + (size_t)(j + 4) * tile_stride + 4096); }
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 0);
      mbar_wait(s_full + 8 * ((j >> 1) & 1), (uint32_t)((j >> 2) & 1));
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 1);
      tc_fence_after();
      float l[64];
      {
        uint32_t raw[32];
        tmem_ld32(tS, raw);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) l[c] = __uint_as_float(raw[c]) * sc[c];
        tmem_ld32(tS + 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) l[32 + c] = __uint_as_float(raw[c]) * sc[32 + c];
      }
      if (ragged && j == T - 1) {
#pragma unroll
        for (int c = 0; c < 64; ++c) l[c] = (j * 64 + c < a.N) ? l[c] : -INFINITY;
      }
      if (j + 2 < T) {  // the SC registers are dead: refill them with this group's next tile under the rest of the work
        load_sc_tile(sc, sc_cta + (size_t)(j + 2) * tile_stride, r);
      }
      float tmax = l[0];
#pragma unroll
      for (int c = 1; c < 64; ++c) tmax = fmaxf(tmax, l[c]);
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 2);
      // running reference maximum of the row, handed from tile to tile between the row's two owner threads
      float prev_ref = -INFINITY;
      if (j > 0) {
        mbar_wait(ref_ready + 8 * ((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));
        prev_ref = ref_s[((j - 1) & 1) * 128 + r];
      }
      const bool advance = (j == 0) || (tmax > prev_ref + kRescaleThreshold);
      const float new_ref = advance ? tmax : prev_ref;
      ref_s[(j & 1) * 128 + r] = new_ref;
      mbar_arrive(ref_ready + 8 * (j & 1));
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 3);
      l_sum *= ex2_approx(my_ref - new_ref);   // the reference may have moved since this thread's previous tile
      my_ref = new_ref;
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        l[c] = ex2_approx(l[c] - new_ref);
        rsum += l[c];
      }
      l_sum += rsum;
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 4);
      // P (16-bit hi / lo images) over this thread's own S row: column c holds keys 2c | 2c+1
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pair<FMT>(l[32 * hf + 2 * i], l[32 * hf + 2 * i + 1], hi[i], lo[i]);
        tmem_st16(tS + 16 * hf, hi);
        if (a.split) tmem_st16(tS + 32 + 16 * hf, lo);
      }
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 5);
      // rare: the reference advanced, O (accumulated under the old reference) must be rescaled before PV_j
      if (__any_sync(0xffffffffu, advance && j > 0)) {
        mbar_wait(pv_done + 8 * ((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));   // PV_{j-1} complete: O quiescent
        tc_fence_after();
        const float scale = (advance && j > 0) ? ex2_approx(prev_ref - new_ref) : 1.0f;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t o[32];
          tmem_ld32(tO + lane_base + c0, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
          tmem_st32(tO + lane_base + c0, o);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + 8 * (j & 3));
      if (stamp) PDSC_STAMP1(a.dbg, j, 1 + g, 6);
    }
    // ---- epilogue: O / l  ->  msg, staged through the (now free) Q region for full-row stores ----
    {
      const int jl = T - 1;
      mbar_wait(ref_ready + 8 * (jl & 1), (uint32_t)((jl >> 1) & 1));
      const float final_ref = ref_s[(jl & 1) * 128 + r];
      lsum_s[g * 128 + r] = l_sum * ex2_approx(my_ref - final_ref);
      if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 3);
      mbar_wait(o_done, 0);   // last PV complete (and with it every earlier MMA)
      tc_fence_after();
      if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 4);
    }
    softmax_all_sync();
    const float inv_l = 1.0f / (lsum_s[r] + lsum_s[128 + r]);
    uint8_t* ostage = smem + kAttnOut;  // [128 rows][512 B], 16-byte chunk c of row r at (c & ~7) | ((c ^ r) & 7)
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {   // group g converts columns [64 g, 64 g + 64) of every row
      uint32_t o[32];
      tmem_ld32(tO + lane_base + 64 * g + c0, o);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = 16 * g + (c0 >> 2) + q;
        *reinterpret_cast<float4*>(ostage + r * 512 + (((c & ~7) | ((c ^ r) & 7)) << 4)) =
            make_float4(__uint_as_float(o[q * 4]) * inv_l, __uint_as_float(o[q * 4 + 1]) * inv_l,
                        __uint_as_float(o[q * 4 + 2]) * inv_l, __uint_as_float(o[q * 4 + 3]) * inv_l);
      }
    }
    tc_fence_before();
    softmax_all_sync();
    float* dst = a.msg + ((size_t)b * a.N + qt * 128) * kC;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int rr = q4 * 32 + 16 * g + i;
      const float4 val = *reinterpret_cast<const float4*>(ostage + rr * 512 + (((lane & ~7) | ((lane ^ rr) & 7)) << 4));
      if (qt * 128 + rr < a.N) *reinterpret_cast<float4*>(dst + (size_t)rr * kC + lane * 4) = val;
    }
    if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 5);
  }
  __syncthreads();
  if (life) PDSC_STAMP1(a.dbg, (blockIdx.x ? 1 : 0), 3, 6);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace pdsc
