// tc_attention: SC-weighted flash attention on tcgen05 (included by encoder_tc.cu only).
//
// Reference: models/PointDSC.py:39-42
//     P = softmax_j( SC_ij * (q_i . k_j) / sqrt(C) ),   msg_i = sum_j P_ij v_j          (heads = 1, C = 128)
// SC multiplies the logit (it is not a mask): SC_ij = 0 leaves logit 0, which still takes softmax mass.
//
// One CTA = 128 queries of one set, looping over 64-key tiles.  Warp roles (320 threads):
//   warp 0      loader : bulk async copies (TMA engine) of the ready-made K / V^T operand images, 2-stage ring
//   warp 1      MMA    : S[j&1] = Q K_j^T (TMEM, double-buffered), O += P_j V_j (TMEM); owns the TMEM allocation
//   warps 2-9   softmax: two warpgroups; thread (row r, half h) owns 32 of the 64 logits of row r.
//               logits (log2 domain; Q carries log2e/sqrt(C)) = S * SC, SC read through its symmetry
//               (SC[j][i], coalesced over the 128 rows); lazily advanced reference maximum (FA4-style): the
//               exponent offset only moves when the row maximum grew by > 8, so O in TMEM is rescaled rarely;
//               P = ex2(l - ref) is split hi/lo into the swizzled smem A-operand image of the PV MMA.
// QK_{j+1} is issued before PV_j, so the tensor core works on the next S tile while the softmax runs.
#pragma once
#include "tc_common.cuh"

namespace pdsc {

struct AttnArgs {
  int N, NS, QT, KT, split;
  const uint8_t* qimg;
  const uint8_t* kvimg;
  const float* sc;
  float* msg;
  long long* dbg;
};

constexpr int kAttnThreads = 320;
constexpr int kAttnQ = 0, kAttnK = 65536, kAttnV = 131072, kAttnP = 196608, kAttnBars = 229376;
constexpr int kAttnMx = kAttnBars + 256;               // float mx[2][2][128]
constexpr int kAttnSmemTc = kAttnMx + 2048;            // 231,680 B (limit 232,448)
constexpr float kRescaleThreshold = 8.0f;              // log2 units: P < 2^8 before the reference max advances

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// streaming read-only load: every SC element is used once per CTA, keep it out of L1
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int FMT>
__global__ void __launch_bounds__(kAttnThreads, 1) tc_attention_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttnBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* mx = reinterpret_cast<float*>(smem + kAttnMx);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t q_full = smem_u32(bars + 0);
  const uint32_t k_full = smem_u32(bars + 1), k_empty = smem_u32(bars + 3);     // [stage] at +8*stage
  const uint32_t v_full = smem_u32(bars + 5), v_empty = smem_u32(bars + 7);
  const uint32_t s_full = smem_u32(bars + 9), s_empty = smem_u32(bars + 11);
  const uint32_t p_full = smem_u32(bars + 13), p_empty = smem_u32(bars + 14);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / a.QT, qt = blockIdx.x % a.QT;
  const int T = a.KT;

  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1);
      mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
      mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 256);
    }
    mbar_init(p_full, 256);
    mbar_init(p_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = tmem + 128;   // S[0] at +0, S[1] at +64 (64 fp32 columns each), O at +128 (128 columns)

  if (warp == 0) {
    // ===================================== loader =====================================
    if (lane == 0) {
      const uint8_t* qsrc = a.qimg + ((size_t)b * a.QT + qt) * 65536;
      mbar_expect_tx(q_full, a.split ? 65536u : 32768u);
      bulk_g2s(s0 + kAttnQ, qsrc, 32768u, q_full);
      if (a.split) bulk_g2s(s0 + kAttnQ + 32768, qsrc + 32768, 32768u, q_full);
      const uint32_t half = a.split ? 32768u : 16384u;
      for (int j = 0; j < T; ++j) {
        const int s = j & 1, u = j >> 1;
        const uint8_t* src = a.kvimg + ((size_t)b * a.KT + j) * 65536;
        if (j >= 2) mbar_wait(k_empty + 8 * s, (uint32_t)((u - 1) & 1));
        mbar_expect_tx(k_full + 8 * s, half);
        bulk_g2s(s0 + kAttnK + s * 32768, src, half, k_full + 8 * s);
        if (j >= 2) mbar_wait(v_empty + 8 * s, (uint32_t)((u - 1) & 1));
        mbar_expect_tx(v_full + 8 * s, half);
        bulk_g2s(s0 + kAttnV + s * 32768, src + 32768, half, v_full + 8 * s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t q_hi = s0 + kAttnQ, q_lo = s0 + kAttnQ + 32768;
      const uint32_t p_hi = s0 + kAttnP, p_lo = s0 + kAttnP + 16384;
      mbar_wait(q_full, 0);
      mbar_wait(k_full, 0);
      tc_fence_after();
      issue_gemm<2, 64>(tmem, q_hi, q_lo, 16384, s0 + kAttnK, s0 + kAttnK + 16384, 8192, a.split, 0, FMT);
      mma_commit(s_full);
      mma_commit(k_empty);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          const int s1 = (j + 1) & 1, u1 = (j + 1) >> 1;
          mbar_wait(k_full + 8 * s1, (uint32_t)(u1 & 1));
          if (j + 1 >= 2) mbar_wait(s_empty + 8 * s1, (uint32_t)((u1 - 1) & 1));
          tc_fence_after();
          const uint32_t kb = s0 + kAttnK + s1 * 32768;
          issue_gemm<2, 64>(tmem + 64 * s1, q_hi, q_lo, 16384, kb, kb + 16384, 8192, a.split, 0, FMT);
          mma_commit(s_full + 8 * s1);
          mma_commit(k_empty + 8 * s1);
        }
        const int s = j & 1, u = j >> 1;
        PDSC_STAMP(a.dbg, j, 0, 0);
        mbar_wait(p_full, (uint32_t)(j & 1));
        PDSC_STAMP(a.dbg, j, 0, 1);
        mbar_wait(v_full + 8 * s, (uint32_t)(u & 1));
        PDSC_STAMP(a.dbg, j, 0, 2);
        tc_fence_after();
        const uint32_t vb = s0 + kAttnV + s * 32768;
        issue_gemm<1, 128>(tO, p_hi, p_lo, 16384, vb, vb + 16384, 16384, a.split, j > 0 ? 1u : 0u, FMT);
        mma_commit(p_empty);
        mma_commit(v_empty + 8 * s);
        PDSC_STAMP(a.dbg, j, 0, 3);
      }
    }
    __syncwarp();
  } else {
    // ===================================== softmax =====================================
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;           // which 32-column half of the 64-key tile this thread owns
    const int r = q4 * 32 + lane;            // query row within the tile == TMEM lane
    const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
    // SC tiles of this CTA: sc_t[b][j][qt][64 keys][128 queries] (sc_matrix.cu); thread (r, h) reads element
    // (32 h + c, r) of tile j at the compile-time offset c * 512 B from one per-tile pointer, coalesced over r.
    const size_t tile_stride = (size_t)a.QT << 13;
    const float* sc_cta = a.sc + ((((size_t)b * a.KT) * a.QT + qt) << 13);
    const float* sc_ptr = sc_cta + (32 * h) * 128 + r;
    const float* sc_line = sc_cta + (tid - 64) * 32;   // one 128-byte line of the 32 KB tile per softmax thread (L2 prefetch)
    uint8_t* Pbuf = smem + kAttnP;
    const bool ragged = (a.N & 63) != 0;
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && tid == 64;
    float m_ref = -INFINITY, l_sum = 0.f;

    float scv[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) scv[c] = ldg_stream(sc_ptr + c * 128);
    if (T > 1) prefetch_l2(sc_line + tile_stride);
    if (T > 2) prefetch_l2(sc_line + 2 * tile_stride);
    for (int j = 0; j < T; ++j) {
      const int s = j & 1, u = j >> 1;
      const int j0 = j * 64 + 32 * h;
      if (j + 3 < T) prefetch_l2(sc_line + (size_t)(j + 3) * tile_stride);
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 0);
      mbar_wait(s_full + 8 * s, (uint32_t)(u & 1));
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 1);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32(tmem + 64 * s + lane_base + 32 * h, raw);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty + 8 * s);

      float p[32];
      float hmax = -INFINITY;
      if (ragged && j == T - 1) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          p[c] = (j0 + c < a.N) ? __uint_as_float(raw[c]) * scv[c] : -INFINITY;
          hmax = fmaxf(hmax, p[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          p[c] = __uint_as_float(raw[c]) * scv[c];
          hmax = fmaxf(hmax, p[c]);
        }
      }
      if (j + 1 < T) {  // the SC registers are dead: refill them with the next tile's values under the rest of this tile
        sc_ptr += tile_stride;
#pragma unroll
        for (int c = 0; c < 32; ++c) scv[c] = ldg_stream(sc_ptr + c * 128);
      }
      // row maximum over both halves
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 2);
      mx[((j & 1) * 2 + h) * 128 + r] = hmax;
      softmax_group_sync();
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 3);
      const float tmax = fmaxf(hmax, mx[((j & 1) * 2 + (1 - h)) * 128 + r]);
      const bool advance = (j == 0) || (tmax > m_ref + kRescaleThreshold);
      const float new_ref = advance ? tmax : m_ref;
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        p[c] = ex2_approx(p[c] - new_ref);
        rsum += p[c];
      }
      const bool rescale_any = __any_sync(0xffffffffu, advance && j > 0);
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 4);
      if (j > 0) {
        mbar_wait(p_empty, (uint32_t)((j - 1) & 1));  // PV_{j-1} done: P smem free, O quiescent
        tc_fence_after();
      }
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 5);
      if (rescale_any) {  // this thread rescales its 64-column half of row r of O
        const float scale = (advance && j > 0) ? ex2_approx(m_ref - new_ref) : 1.0f;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t o[32];
          tmem_ld32(tO + lane_base + 64 * h + c0, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
          tmem_st32(tO + lane_base + 64 * h + c0, o);
        }
        tmem_st_wait();
        l_sum *= scale;
      }
      m_ref = new_ref;
      l_sum += rsum;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8<FMT>(p + g * 8, hi, lo);
        const uint32_t off = sw128_offset((uint32_t)r, (uint32_t)(32 * h + g * 8));
        *reinterpret_cast<uint4*>(Pbuf + off) = hi;
        if (a.split) *reinterpret_cast<uint4*>(Pbuf + 16384 + off) = lo;
      }
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 6);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (stamp) PDSC_STAMP1(a.dbg, j, 1, 7);
    }
    // ---- epilogue: O / l  ->  msg, staged through the (now free) Q region for full-row stores ----
    const int lb = (T & 1) * 2;  // the mx buffer NOT used by tile T-1 (its last readers are behind tile T-1's group sync)
    mx[(lb + h) * 128 + r] = l_sum;
    mbar_wait(p_empty, (uint32_t)((T - 1) & 1));
    tc_fence_after();
    softmax_group_sync();
    const float inv_l = 1.0f / (l_sum + mx[(lb + 1 - h) * 128 + r]);
    uint8_t* ostage = smem + kAttnQ;  // [128 rows][512 B], 16-byte chunk c of row r at (c & ~7) | ((c ^ r) & 7)
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(tO + lane_base + 64 * h + c0, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int c = 16 * h + (c0 >> 2) + g;
        *reinterpret_cast<float4*>(ostage + r * 512 + (((c & ~7) | ((c ^ r) & 7)) << 4)) =
            make_float4(__uint_as_float(o[g * 4]) * inv_l, __uint_as_float(o[g * 4 + 1]) * inv_l,
                        __uint_as_float(o[g * 4 + 2]) * inv_l, __uint_as_float(o[g * 4 + 3]) * inv_l);
      }
    }
    tc_fence_before();
    softmax_group_sync();
    float* dst = a.msg + ((size_t)b * a.N + qt * 128) * kC;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int rr = q4 * 32 + 16 * h + i;
      const float4 val = *reinterpret_cast<const float4*>(ostage + rr * 512 + (((lane & ~7) | ((lane ^ rr) & 7)) << 4));
      if (qt * 128 + rr < a.N) *reinterpret_cast<float4*>(dst + (size_t)rr * kC + lane * 4) = val;
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace pdsc
