// tc_attention (persistent): SC-weighted flash attention on tcgen05, one CTA per SM looping over (set, 128-query tile)
// work items (included by encoder_tc.cu only).
//
// Reference: models/PointDSC.py:39-42
//     P = softmax_j( SC_ij * (q_i . k_j) / sqrt(C) ),   msg_i = sum_j P_ij v_j          (heads = 1, C = 128)
// SC multiplies the logit (it is not a mask): SC_ij = 0 leaves logit 0, which still takes softmax mass.
//
// The per-item machinery is that of tc_attention.cuh (Q and P as A operands in tensor memory, two softmax groups on
// alternate key tiles with the running row maximum handed over through shared memory, lazy rescale of O), except that
// S = Q K^T is issued and committed per 64-key tile (N = 64: the tensor time is the same as for N = 128 pairs, but the two
// groups receive their tiles at different times and stay staggered instead of contending for the same pipes in the same
// phase).  What is new is that NOTHING is torn down between items: barriers keep their phase (running use counts or
// explicit phase bits), the K / V rings and the S / P buffers keep rotating, and
//   * the loader streams the next item's K / V tiles and its Q image (through a 32 KB staging buffer, hi then lo half)
//     while the current item is still being computed;
//   * the softmax groups move the next Q into tensor memory right after their last tile of the current item (the Q
//     columns are free once the item's last QK pair has completed);
//   * the MMA warp issues the next item's first four QK tiles directly behind the current item's last PV, so the tensor
//     core works on them while the softmax groups drain O;
//   * in the MMA warp every normally-satisfied wait (K pair, V tile, O drained) is taken before the wait for P_j, so one
//     mbarrier wake-up separates a group's arrival from the PV_j / QK_{j+4} issue (a satisfied wait still costs
//     200-300 cycles on a sub-partition shared with two busy softmax warps).
// Per-item fixed cost drops from ~14 k cycles (TMEM allocation, barrier set-up, cold Q / K / SC loads, output drain,
// CTA launch) to the ~2 k-cycle output drain.
#pragma once
#include "tc_attention.cuh"

namespace pdsc {

constexpr int kAttnPK = 0;                                   // K: 3 tile stages x 32 KB ([hi p0][hi p1][lo p0][lo p1], 8 KB each)
constexpr int kAttnPV = 98304;                               // V: 3 tile stages x 32 KB
constexpr int kAttnPQ = 196608;                              // Q staging: one 32 KB half image
constexpr int kAttnPBars = 229376;
constexpr int kAttnPRef = kAttnPBars + 320;                  // float ref[2][128], lsum[2][128]
constexpr int kAttnPSmem = kAttnPRef + 2048;                 // 231,744 B (limit 232,448)
constexpr int kAttnRing = 3;                                 // K / V ring depth: a stage freed by MMA t is refilled for tile t + 3

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); }

template <int FMT>
__global__ void __launch_bounds__(kAttnThreads, 1) tc_attention_persistent_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttnPBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);
  float* ref_s = reinterpret_cast<float*>(smem + kAttnPRef);  // [2][128] reference maximum after tile j (slot j & 1)
  float* lsum_s = ref_s + 256;                                // [2][128] per-group row sums (epilogue)
  const uint32_t s0 = smem_u32(smem);
  const uint32_t k_full = smem_u32(bars + 0), k_empty = smem_u32(bars + 3);      // [3] per K tile stage
  const uint32_t v_full = smem_u32(bars + 6), v_empty = smem_u32(bars + 9);      // [3] per V tile stage
  const uint32_t p_full = smem_u32(bars + 12), s_full = smem_u32(bars + 16);     // [4] per S/P tile buffer
  const uint32_t pv_done = smem_u32(bars + 20), ref_ready = smem_u32(bars + 22); // [2]
  const uint32_t qh_full = smem_u32(bars + 24), ql_full = smem_u32(bars + 25);   // staging holds the hi / lo half image
  const uint32_t qh_used = smem_u32(bars + 26), ql_used = smem_u32(bars + 27);   // ... and has been moved to TMEM
  const uint32_t q_tmem = smem_u32(bars + 28), o_done = smem_u32(bars + 29), o_free = smem_u32(bars + 30);
  const uint32_t stage_free = smem_u32(bars + 31);   // the item's output has left the staging buffer
  const uint32_t qk_done = smem_u32(bars + 32);      // the item's last QK has completed: the Q columns are free
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.TS;                                          // key tiles per work item (== a.KT unless the keys are split)
  const int TE = (T + 1) >> 1, TO = T >> 1;                   // tiles per item with even / odd index
  const int my_items = (a.items > (int)blockIdx.x) ? (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  const long long t_begin = (a.dbg != nullptr) ? clock64() : 0;
  unsigned long long ns_begin = 0;
  if (a.dbg != nullptr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < kAttnRing; ++i) {
      mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1);
      mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(pv_done + 8 * i, 1); mbar_init(ref_ready + 8 * i, 128); }
    for (int i = 0; i < 4; ++i) { mbar_init(p_full + 8 * i, 128); mbar_init(s_full + 8 * i, 1); }
    mbar_init(qh_full, 1); mbar_init(ql_full, 1);
    mbar_init(qh_used, 128); mbar_init(ql_used, 128);
    mbar_init(qk_done, 1);
    mbar_init(q_tmem, 256); mbar_init(o_done, 1); mbar_init(o_free, 256); mbar_init(stage_free, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = tmem + 256;   // S/P tile buffer i at +64 i, O at +256 (128 fp32 columns)
  const uint32_t tQ = tmem + 384;   // Q: hi image (64 columns = 128 channels) at +384, lo image at +448

  if (warp == 0) {
    // ===================================== loader =====================================
    if (lane == 0) {
      // three independent streams over this CTA's items, each gated only by its own ring / staging slot
      int ki = 0, kj = 0, kst = 0, kuse = 0;   // K: item ordinal, tile within the item, ring stage and its use count
      int vi = 0, vj = 0, vst = 0, vuse = 0;   // V: the same
      int qi = 0, qh = 0;                      // Q: item ordinal, half (0 = hi, 1 = lo)
      const int qhalves = a.split ? 2 : 1;
      const uint32_t tile_bytes = a.split ? 32768u : 16384u;   // hi (+ lo) image of one 64-key tile
      while (ki < my_items || vi < my_items || qi < my_items) {
        bool progress = false;
        if (ki < my_items) {
          if (kuse == 0 || mbar_test(k_empty + 8 * kst, (uint32_t)((kuse - 1) & 1))) {
            const int witem = blockIdx.x + ki * gridDim.x;
            const int item = witem / a.splits;                                       // (set, query tile)
            const int kt = min((witem % a.splits) * T + kj, a.KT - 1);               // a virtual tile re-reads the last real one (it is masked)
            const uint8_t* kv = a.kvimg + ((size_t)(item / a.QT) * a.KT + kt) * 65536;
            mbar_expect_tx(k_full + 8 * kst, tile_bytes);
            bulk_g2s(s0 + kAttnPK + kst * 32768, kv, tile_bytes, k_full + 8 * kst);
            if (++kst == kAttnRing) { kst = 0; ++kuse; }
            if (++kj == T) { kj = 0; ++ki; }
            progress = true;
          }
        }
        if (vi < my_items) {
          if (vuse == 0 || mbar_test(v_empty + 8 * vst, (uint32_t)((vuse - 1) & 1))) {
            const int witem = blockIdx.x + vi * gridDim.x;
            const int item = witem / a.splits;
            const int kt = min((witem % a.splits) * T + vj, a.KT - 1);
            const uint8_t* kv = a.kvimg + ((size_t)(item / a.QT) * a.KT + kt) * 65536 + 32768;
            mbar_expect_tx(v_full + 8 * vst, tile_bytes);
            bulk_g2s(s0 + kAttnPV + vst * 32768, kv, tile_bytes, v_full + 8 * vst);
            if (++vst == kAttnRing) { vst = 0; ++vuse; }
            if (++vj == T) { vj = 0; ++vi; }
            progress = true;
          }
        }
        if (qi < my_items) {
          // the staging buffer is free once the previous half has been moved to tensor memory
          bool free;
          // (the buffer also stages the output tile of item qi - 2, after that item's Q(qi - 1) halves have gone through)
          if (qh == 0) free = (qi == 0) || (mbar_test(a.split ? ql_used : qh_used, (uint32_t)((qi - 1) & 1)) &&
                                            (qi < 2 || mbar_test(stage_free, (uint32_t)((qi - 2) & 1))));
          else free = mbar_test(qh_used, (uint32_t)(qi & 1));
          if (free) {
            const int item = (blockIdx.x + qi * gridDim.x) / a.splits;
            const uint8_t* qsrc = a.qimg + (size_t)item * 65536 + qh * 32768;
            const uint32_t bar = qh ? ql_full : qh_full;
            mbar_expect_tx(bar, 32768u);
            bulk_g2s(s0 + kAttnPQ, qsrc, 32768u, bar);
            if (qh == 0 && a.split)   // pull the lo half towards L2 now: it is staged late, at the item boundary
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qsrc + 32768), "r"(32768u) : "memory");
            if (++qh == qhalves) { qh = 0; ++qi; }
            progress = true;
          }
        }
        if (!progress) __nanosleep(400);   // the rings turn over every ~2.5k cycles: poll rarely, the issue slots are the softmax warps'
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    const bool leader = elect_one();
    // S_t = Q K_t^T for ONE 64-key tile (N = 64, 24 MMAs of 32 cycles) into S/P buffer t & 3, t the CTA's running tile
    // count.  One commit per tile: the two softmax groups receive their tiles 768 tensor cycles apart and stay staggered
    // (with N = 128 pairs both groups ran the same phase at the same time and contended for the same pipes).
    // QK runs kAttnRing tiles ahead of PV: behind PV_j, tile j + 3 overwrites the S/P buffer that PV_{j-1} has read (its
    // completion is vouched for by the softmax group before it arrives on p_full(j)), from the K stage QK_j released.
    int qst = 0;                                   // K ring stage of the next QK
    auto issue_qk_tile = [&](int gt, bool last) {  // gt: running tile count; last: the item's final tile
      const int buf = gt & 3;
      if (leader) {
        const uint32_t kb = s0 + kAttnPK + qst * 32768;
        issue_gemm_ts<2, 64>(tmem + 64 * buf, tQ, tQ + 64, kb, kb + 16384, 8192, a.split, 0, FMT);
        mma_commit(s_full + 8 * buf);
        mma_commit(k_empty + 8 * qst);
        if (last) mma_commit(qk_done);
      }
      if (++qst == kAttnRing) qst = 0;
    };
    int gv = 0, vst = 0;
    const bool stamp_mma = leader && a.dbg != nullptr && blockIdx.x == 0;
    for (int it = 0; it < my_items; ++it) {
      const int gtb = it * T;
      // the item's first tiles: K resident (the later ones are vouched for by the softmax groups, see below)
      for (int j = 0; j < T && j < kAttnRing; ++j) {
        const int g3 = gtb + j;
        mbar_wait(k_full + 8 * (g3 % kAttnRing), (uint32_t)((g3 / kAttnRing) & 1));
      }
      mbar_wait(q_tmem, (uint32_t)(it & 1));   // this item's Q (hi | lo images) is resident in tensor memory
      tc_fence_after();
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 0);
      for (int j = 0; j < T && j < kAttnRing; ++j) issue_qk_tile(gtb + j, j == T - 1);
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 1);
      for (int j = 0; j < T; ++j, ++gv) {
        const int buf = gv & 3;
        // p_full(j) also vouches for V_j and for K_{j+3}: the softmax group waits for those (normally long-satisfied)
        // barriers before it arrives, so this warp - the critical path of the kernel, and slow per instruction on a
        // sub-partition it shares with two busy softmax warps - has ONE wait per tile
        if (j == 0 && it > 0) mbar_wait(o_free, (uint32_t)((it - 1) & 1));   // previous item's O has been drained
        mbar_wait(p_full + 8 * buf, (uint32_t)((gv >> 2) & 1));
        tc_fence_after();
        if (stamp_mma && it == 2) PDSC_STAMP1(a.dbg, j, 3, 0);   // P_j seen
        if (leader) {
          const uint32_t vb = s0 + kAttnPV + vst * 32768;
          const uint32_t tP = tmem + 64 * buf;   // P_j: hi image in columns [0,32), lo image in [32,64) of its S tile
          issue_pv_mn(tO, tP, tP + 32, vb, vb + 16384, a.split, j > 0 ? 1u : 0u, FMT);
          mma_commit(pv_done + 8 * (gv & 1));
          mma_commit(v_empty + 8 * vst);
        }
        if (++vst == kAttnRing) vst = 0;
        if (stamp_mma && it == 2) PDSC_STAMP1(a.dbg, j, 3, 7);   // PV_j issued
        if (stamp_mma && j == 0) PDSC_STAMP1(a.dbg, it, 0, 2);
        if (stamp_mma && j == T - 1) PDSC_STAMP1(a.dbg, it, 0, 3);
        if (j + kAttnRing < T) issue_qk_tile(gv + kAttnRing, j + kAttnRing == T - 1);
        if (stamp_mma && it == 2) PDSC_STAMP1(a.dbg, j, 3, 5);   // QK_{j+3} issued
      }
      if (leader) mma_commit(o_done);
    }
    __syncwarp();
  } else {
    // ===================================== softmax =====================================
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int g = (warp - 2) >> 2;           // group: tiles j with (j & 1) == g
    const int r = q4 * 32 + lane;            // query row within the tile == TMEM lane
    const int gt = (warp - 2 - 4 * g) * 32 + lane;   // thread index within the group
    const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
    const size_t tile_stride = (size_t)a.QT << 13;
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && gt == 0;

    // move one half of the staged Q image (this thread's row) into tensor memory: hi by group 0, lo by group 1
    auto convert_q = [&](int it_next) {
      if (g == 0 || a.split) {
        mbar_wait(g ? ql_full : qh_full, (uint32_t)(it_next & 1));
        const uint8_t* qrow = smem + kAttnPQ + (r >> 3) * 1024 + (r & 7) * 128;
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
        mbar_arrive(g ? ql_used : qh_used);
      }
      tc_fence_before();
      mbar_arrive(q_tmem);
    };

    float sc[64];
    auto load_sc = [&](int witem, int j) {   // this thread's 64 SC values of tile j of work item `witem`: 16 coalesced float4 loads
      const int item = witem / a.splits;
      const int kt = min((witem % a.splits) * T + j, a.KT - 1);
      load_sc_tile(sc, a.sc + ((((size_t)(item / a.QT) * a.KT) * a.QT + (item % a.QT)) << 13) + (size_t)kt * tile_stride, r);
    };
    if (my_items > 0) {
      if (g < T) load_sc(blockIdx.x, g);
      convert_q(0);
    }
    for (int it = 0; it < my_items; ++it) {
      const int witem = blockIdx.x + it * gridDim.x;
      const int item = witem / a.splits, t0 = (witem % a.splits) * T;   // (set, query tile) and the split's first key tile
      const int b = item / a.QT, qt = item % a.QT;
      const int gvb = it * T;   // the CTA's running tile count at the item's first tile
      const float* sc_cta = a.sc + ((((size_t)b * a.KT) * a.QT + qt) << 13);
      const float* sc_line = sc_cta + gt * 32;   // two 128-byte lines of each 32 KB tile per thread (L2 prefetch)
      float my_ref = -INFINITY, l_sum = 0.f;
      if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 0);
      if (g + 2 < T && t0 + g + 2 < a.KT) { prefetch_l2(sc_line + (size_t)(t0 + g + 2) * tile_stride); prefetch_l2(sc_line + (size_t)(t0 + g + 2) * tile_stride + 4096); }
      for (int j = g; j < T; j += 2) {
        const int tn = gvb + j;   // running tile count
        const int buf = tn & 3;
        const uint32_t tS = tmem + 64 * buf + lane_base;
        if (j + 4 < T && t0 + j + 4 < a.KT) { prefetch_l2(sc_line + (size_t)(t0 + j + 4) * tile_stride); prefetch_l2(sc_line + (size_t)(t0 + j + 4) * tile_stride + 4096); }
        mbar_wait(s_full + 8 * buf, (uint32_t)((tn >> 2) & 1));
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
        if ((t0 + j) * 64 + 63 >= a.N) {     // the set's last, ragged key tile - or a virtual tile behind it (all masked)
#pragma unroll
          for (int c = 0; c < 64; ++c) l[c] = ((t0 + j) * 64 + c < a.N) ? l[c] : -INFINITY;
        }
        // the SC registers are dead: refill them with this group's next tile (of this item or of the next one)
        if (j + 2 < T) load_sc(witem, j + 2);
        else if (it + 1 < my_items && g < T) load_sc(witem + gridDim.x, g);
        float tmax = l[0];
#pragma unroll
        for (int c = 1; c < 64; ++c) tmax = fmaxf(tmax, l[c]);
        // running reference maximum of the row, handed from tile to tile between the row's two owner threads
        float prev_ref = -INFINITY;
        if (j > 0) {
          const int pj = j - 1;
          const int use = ((pj & 1) ? it * TO : it * TE) + (pj >> 1);
          mbar_wait(ref_ready + 8 * (pj & 1), (uint32_t)(use & 1));
          prev_ref = ref_s[(pj & 1) * 128 + r];
        }
        const bool advance = (j == 0) || (tmax > prev_ref + kRescaleThreshold);
        const float new_ref = advance ? tmax : prev_ref;
        ref_s[(j & 1) * 128 + r] = new_ref;
        mbar_arrive(ref_ready + 8 * (j & 1));
        l_sum *= ex2_approx(my_ref - new_ref);   // the reference may have moved since this thread's previous tile
        my_ref = new_ref;
        float rsum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          l[c] = ex2_approx(l[c] - new_ref);
          rsum += l[c];
        }
        l_sum += rsum;
        // P (16-bit hi / lo images) over this thread's own S row: column c holds keys 2c | 2c+1
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) split_pair<FMT>(l[32 * hf + 2 * i], l[32 * hf + 2 * i + 1], hi[i], lo[i]);
          tmem_st16(tS + 16 * hf, hi);
          if (a.split) tmem_st16(tS + 32 + 16 * hf, lo);
        }
        // rare: the reference advanced, O (accumulated under the old reference) must be rescaled before PV_j
#ifdef PDSC_EXP_NO_RESCALE          // timing experiment only (tools/build_variant.py): results are wrong
        if (false) {
#else
        if (__any_sync(0xffffffffu, advance && j > 0)) {
#endif
          const int gp = gvb + j - 1;
          mbar_wait(pv_done + 8 * (gp & 1), (uint32_t)((gp >> 1) & 1));   // PV_{j-1} complete: O quiescent
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
        // vouch for the operands the MMA warp will issue behind P_j: V_j and K_{j+3} (loaded three tile periods ago)
        {
          // ... and for the COMPLETION of PV_{j-1}: QK_{j+3}, issued behind PV_j, overwrites the S/P buffer PV_{j-1} reads as
          // its A operand (rule in tc_common.cuh).  For j = 0 the o_free wait of the MMA warp covers the previous item's PVs.
          if (j > 0) mbar_wait(pv_done + 8 * ((tn - 1) & 1), (uint32_t)(((tn - 1) >> 1) & 1));
          mbar_wait(v_full + 8 * (tn % kAttnRing), (uint32_t)((tn / kAttnRing) & 1));
          if (j + kAttnRing < T) {
            const int gk3 = tn + kAttnRing;
            mbar_wait(k_full + 8 * (gk3 % kAttnRing), (uint32_t)((gk3 / kAttnRing) & 1));
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full + 8 * buf);
      }
      // ---- item boundary ------------------------------------------------------------------------------------
      if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 2);
      if (it + 1 < my_items) {
        mbar_wait(qk_done, (uint32_t)(it & 1));   // the item's last QK has completed: the Q columns are free
        tc_fence_after();
        if (g >= T) load_sc(witem + gridDim.x, g);   // (never true for T >= 2; keeps a one-tile item correct)
        convert_q(it + 1);
      }
      if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 3);
      // ---- epilogue: O / l -> msg, straight from registers (the shared memory belongs to the next item's tiles) ----
      {
        const int jl = T - 1;
        const int use = ((jl & 1) ? it * TO : it * TE) + (jl >> 1);
        mbar_wait(ref_ready + 8 * (jl & 1), (uint32_t)(use & 1));
        const float final_ref = ref_s[(jl & 1) * 128 + r];
        lsum_s[g * 128 + r] = l_sum * ex2_approx(my_ref - final_ref);
        mbar_wait(o_done, (uint32_t)(it & 1));   // last PV complete (and with it every earlier MMA)
        tc_fence_after();
        if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 4);
      }
      softmax_all_sync();
      if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 5);
      // one query tile per item: msg = O / l.  Key split: the item's UNNORMALISED O, its reference maximum and its row sum go
      // to the partial buffers; tc_attention_merge_kernel combines the splits of a query tile in ascending split order.
      const bool partial = a.splits > 1;
      const float l_tot = lsum_s[r] + lsum_s[128 + r];
      const float inv_l = partial ? 1.0f : 1.0f / l_tot;
      if (partial && g == 0) {
        const int jlp = T - 1;
        *reinterpret_cast<float2*>(a.part_ml + ((size_t)witem * 128 + r) * 2) = make_float2(ref_s[(jlp & 1) * 128 + r], l_tot);
      }
      // group g drains columns [64 g, 64 g + 64) of every row, 32 columns at a time, through its 16 KB half of the Q
      // staging buffer ([128 rows][128 B], 16-byte chunks XOR-swizzled by row) so that every store instruction writes
      // four full 128-byte lines.  The buffer is free here: both halves of the next Q have already gone through it.
      // (split formats: group 1 moves the lo image, which fills the whole buffer, so group 0 waits for it; single formats: only
      // group 0 moves an image - the hi one, also the whole buffer - so group 1 waits for that)
      if (it + 1 < my_items) {
        if (a.split && g == 0) mbar_wait(ql_used, (uint32_t)((it + 1) & 1));
        if (!a.split && g == 1) mbar_wait(qh_used, (uint32_t)((it + 1) & 1));
      }
      uint8_t* ost = smem + kAttnPQ + g * 16384;
      float* dst = partial ? a.part_o + (size_t)witem * (128 * kC) + 64 * g : a.msg + ((size_t)b * a.N + qt * 128) * kC + 64 * g;
      const int rsub = gt >> 3, piece = gt & 7;   // read-out: rows rsub + 16 i, 16-byte piece of the 128-byte segment
#pragma unroll
      for (int sr = 0; sr < 2; ++sr) {
        uint32_t o[32];
        tmem_ld32(tO + lane_base + 64 * g + 32 * sr, o);
        tmem_ld_wait();
        if (sr == 1) {
          tc_fence_before();
          mbar_arrive(o_free);                  // O may be overwritten by the next item's first PV
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(ost + r * 128 + ((q ^ (r & 7)) << 4)) =
              make_float4(__uint_as_float(o[4 * q]) * inv_l, __uint_as_float(o[4 * q + 1]) * inv_l,
                          __uint_as_float(o[4 * q + 2]) * inv_l, __uint_as_float(o[4 * q + 3]) * inv_l);
        group_sync(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = rsub + 16 * i;
          const float4 val = *reinterpret_cast<const float4*>(ost + rr * 128 + ((piece ^ (rr & 7)) << 4));
          if (partial || qt * 128 + rr < a.N) *reinterpret_cast<float4*>(dst + (size_t)rr * kC + 32 * sr + piece * 4) = val;
        }
        group_sync(g);
      }
      mbar_arrive(stage_free);
      if (stamp) PDSC_STAMP1(a.dbg, it, 1 + g, 6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (a.dbg != nullptr && tid == 0) {   // spread of the CTA lifetimes: rows 14 / 15 of role 0 (never used by an item)
    const long long dt = clock64() - t_begin;
    atomicMax(reinterpret_cast<unsigned long long*>(a.dbg + (14 * 4) * 8 + 0), (unsigned long long)dt);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + (14 * 4) * 8 + 1), (unsigned long long)dt);
    if (my_items == 14) atomicMax(reinterpret_cast<unsigned long long*>(a.dbg + (14 * 4) * 8 + 2), (unsigned long long)dt);
    if (blockIdx.x < 8) a.dbg[(15 * 4) * 8 + blockIdx.x] = dt;
    if (blockIdx.x == 0) {   // SM clock during this launch = cycles / ns
      unsigned long long ns_end;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
      a.dbg[(15 * 4 + 1) * 8 + 0] = (long long)(ns_end - ns_begin);
      a.dbg[(15 * 4 + 1) * 8 + 1] = dt;
    }
  }
}

// ---- key split: combine the partial results of one query tile -------------------------------------------------------------
// msg_i = sum_s O_s[i] 2^(m_s - m*) / sum_s l_s 2^(m_s - m*),  m* = max_s m_s, splits added in ascending order (deterministic).
// One CTA per (query tile, 32-row quarter), thread = (row, 16-byte column piece stride).
__global__ void __launch_bounds__(256) tc_attention_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                                 float* __restrict__ msg, int N, int QT, int splits) {
  const int item = blockIdx.x >> 2, quarter = blockIdx.x & 3;
  const int b = item / QT, qt = item % QT;
  const int row = quarter * 32 + (threadIdx.x >> 3);
  if (qt * 128 + row >= N) return;
  // the reference maxima first (independent loads), then the partial rows four splits at a time so that their loads overlap:
  // at bs = 1 this kernel is pure L2 latency (8 splits x 5 dependent round trips took 9.5 us per layer)
  float mstar = -INFINITY;
#pragma unroll 4
  for (int s = 0; s < splits; ++s) mstar = fmaxf(mstar, part_ml[(((size_t)item * splits + s) * 128 + row) * 2]);
  float L = 0.f;
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < splits; s0 += 4) {
    float2 ml[4];
    float4 v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + u < splits ? s0 + u : splits - 1;      // a clamped re-read of the last split is given weight 0 below
      const size_t w = (size_t)item * splits + s;
      ml[u] = *reinterpret_cast<const float2*>(part_ml + (w * 128 + row) * 2);
      const float* o = part_o + (w * 128 + row) * kC;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[u][i] = *reinterpret_cast<const float4*>(o + ((threadIdx.x & 7) + 8 * i) * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {       // ascending split order: the sum is the same as one split at a time
      if (s0 + u < splits) {
        const float wgt = ex2_approx(ml[u].x - mstar);
        L = fmaf(ml[u].y, wgt, L);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x = fmaf(v[u][i].x, wgt, acc[i].x); acc[i].y = fmaf(v[u][i].y, wgt, acc[i].y);
          acc[i].z = fmaf(v[u][i].z, wgt, acc[i].z); acc[i].w = fmaf(v[u][i].w, wgt, acc[i].w);
        }
      }
    }
  }
  const float inv = 1.0f / L;
  float* dst = msg + ((size_t)b * N + qt * 128 + row) * kC;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + ((threadIdx.x & 7) + 8 * i) * 4) = make_float4(acc[i].x * inv, acc[i].y * inv, acc[i].z * inv, acc[i].w * inv);
}

}  // namespace pdsc
