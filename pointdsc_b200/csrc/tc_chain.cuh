// tc_chain: fused row-tile GEMM chains of the encoder's 1x1 convolutions on tcgen05 (included by encoder_tc.cu).
//
//   PCQ : feat  --W1,relu--> feat1 (fp32 -> HBM) --Wq--> Q image (HBM)
//   KV  : feat1 --Wk--> K image (HBM) ;  feat1 --Wv--> V image (HBM, same row-major format as K)
//   MSG : msg --Wm0,relu--> --Wm1,relu--> --Wm2--> + feat1 --> feat (fp32 -> HBM)
// (reference models/PointDSC.py:56-61 PointCN, :21-23/:36-38 projections, :12-20/:43-44 fc_message + residual)
//
// Persistent CTAs (one per SM), weights resident in shared memory, 128-row tiles.  Warp roles (512 threads, so that the
// register file divides into 128 registers per thread):
//   warps 0-7   epilogue: two warpgroups; thread (row r, half h) owns TMEM lane r and half of the step's columns;
//               bias / ReLU / residual, hi-lo split, stores
//   warps 8-14  loaders : prefetch the NEXT tile's fp32 rows into registers (coalesced, rows lw, lw+7, ...), convert
//               to the swizzled 16-bit A image once the tensor core has released the buffer
//   warp  15    MMA issuer (whole warp runs the control flow, one elected lane issues) + TMEM allocation
// Chained steps never go back through shared memory: the epilogue writes the next step's A operand (16-bit hi | lo
// images) over the accumulator columns it has just read, and the next MMA takes A FROM TENSOR MEMORY.  Chunk q (K
// elements 32 q .. 32 q + 31) of such an operand sits at columns 32 q .. 32 q + 31 of the producing accumulator as
// [hi: 16 columns | lo: 16 columns].  The shared-memory A buffer is therefore released as soon as the tile's first
// MMA has consumed it, and the loaders convert tile t+1 under the epilogue of tile t.
// Accumulators are double-buffered in TMEM by tile parity (2 x 256 columns).  Global stores are staged through a
// per-warp swizzled smem buffer so that every store instruction writes full 64/128-byte segments.
#pragma once
#include "tc_common.cuh"

namespace pdsc {

constexpr int kChainThreads = 512;
constexpr int kChLoaderWarps = 7, kChLoaderRows = 19;   // rows lw + 7 i, i < 19 (the last one only for lw < 2)
constexpr int kChA = 0;                          // A image: [hi p0 16K][hi p1 16K][lo p0 16K][lo p1 16K]
constexpr int kChW = 65536;                      // weight images (128 KB for PCQ / KV, 80 KB for MSG)
constexpr int kChStage = 65536 + 131072;         // PCQ / KV: 8 x 4 KB store staging
constexpr int kChRes = 65536 + 81920;            // MSG: 64 KB residual tile (also the fp32 store staging)
constexpr int kChBias = kChStage + 32768;        // 256 floats: this mode's biases
constexpr int kChBars = kChBias + 1024;
constexpr int kChainSmem = kChBars + 256;        // 230,656 B
static_assert(kChRes + 65536 <= kChBias, "smem map");

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// explicit global-space 16-byte store: through a pointer the compiler cannot prove global (an element of a local
// pointer array, a select with nullptr) it emits generic ST.E, which is markedly slower than STG
__device__ __forceinline__ void st_global_v4(void* p, const uint4& v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_global_u16(void* p, uint16_t v) {
  asm volatile("st.global.b16 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "h"(v) : "memory");
}
__device__ __forceinline__ void quarter_sync(int q4) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q4) : "memory"); }

// D[128 x NOUT] (+)= A[128 x 32 KCH] * B[NOUT x 32 KCH]^T, A in tensor memory in the chunked in-place layout described
// above (chunk q at a_base + 32 q), B K-major SWIZZLE_128B panels of 64 K elements in shared memory.
template <int KCH, int NOUT>
__device__ __forceinline__ void issue_gemm_tchunk(uint32_t d_tmem, uint32_t a_base, uint32_t b_hi, uint32_t b_lo,
                                                  uint32_t b_panel_bytes, int split, int fmt) {
  const uint32_t idesc = idesc_f16kind(128, NOUT, fmt);
  constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t > 0 && !split) break;
    const uint32_t b = (t == 1) ? b_lo : b_hi;
#pragma unroll
    for (int q = 0; q < KCH; ++q) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int k0 = 32 * q + 16 * s;
        const uint32_t acol = a_base + 32 * q + (t == 2 ? 16 : 0) + 8 * s;
        const uint32_t blo = (((b + (k0 >> 6) * b_panel_bytes + ((k0 & 63) >> 4) * 32) >> 4) & 0x3FFFu) | (1u << 16);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(acol), "r"(blo), "r"(idesc), "r"(acc), "r"(kDescHi)
            : "memory");
        acc = 1;
      }
    }
  }
}

// (set, index within the set) of the row `rr` rows after a row known to be (b0, n0); rows < 2^31
__device__ __forceinline__ void locate_row(int b0, int n0, int rr, int N, int& bb, int& nn) {
  nn = n0 + rr;
  bb = b0;
  if (N >= 64) {            // at most one wrap within a 32-row quarter
    if (nn >= N) { nn -= N; ++bb; }
  } else {
    bb += nn / N;
    nn = nn % N;
  }
}

template <int MODE, int FMT>
__global__ void __launch_bounds__(kChainThreads, 1) tc_chain_kernel(ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Abuf = smem + kChA;
  float* bias = reinterpret_cast<float*>(smem + kChBias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bar_w = smem_u32(bars + 0), a_ready = smem_u32(bars + 1), a1_ready = smem_u32(bars + 2),
                 a_free = smem_u32(bars + 3), r_ready = smem_u32(bars + 4), r_free = smem_u32(bars + 5);
  const uint32_t d_full = smem_u32(bars + 6);   // [step][parity] at + 8 * (step * 2 + parity)
  const uint32_t d_free = smem_u32(bars + 12);  // [parity]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_base = s0 + kChA, w_base = s0 + kChW;
  constexpr int kSteps = (MODE == kMSG) ? 3 : 2;
  // this mode's biases, packed: PCQ b1|bq, KV bk|bv, MSG bm0|bm1|bm2
  constexpr int kBiasSrc = (MODE == kPCQ) ? kB1 : (MODE == kKV) ? kBk : kBm0;

  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(bar_w, 1);
    mbar_init(a_ready, kChLoaderWarps * 32);
    mbar_init(a1_ready, 256);
    mbar_init(a_free, 1);
    mbar_init(r_ready, kChLoaderWarps * 32);
    mbar_init(r_free, 256);
    for (int i = 0; i < 6; ++i) mbar_init(d_full + 8 * i, 1);
    mbar_init(d_free, 256);
    mbar_init(d_free + 8, 256);
    fence_barrier_init();
  }
  if (warp == 15) tmem_alloc(smem_u32(tmem_slot), 512);
  if (tid < 256) bias[tid] = a.bias[kBiasSrc + tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_w, (uint32_t)a.wbytes);
    for (int off = 0; off < a.wbytes; off += 32768)
      bulk_g2s(w_base + off, a.wimg + off, (uint32_t)min(32768, a.wbytes - off), bar_w);
  }
  const long long num_tiles = (a.rows + 127) / 128;

  if (warp == 15) {
    // =================================== MMA issuer ===================================
    const bool leader = elect_one();
    const bool stamp_mma = leader && a.dbg != nullptr && blockIdx.x == 0;
    mbar_wait(bar_w, 0);
    int it = 0;
    uint32_t a1_uses = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int par = it & 1, u = it >> 1;
      const uint32_t dcol = tmem + (uint32_t)par * 256u;
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 0);
      mbar_wait(a_ready, (uint32_t)(it & 1));
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 1);
      // PCQ: D0[par] is known drained (the a1_ready wait of tile t-2 covered it); D1[par] is checked before step 1
      if (MODE != kPCQ && it >= 2) mbar_wait(d_free + 8 * par, (uint32_t)((u - 1) & 1));  // epilogue drained D[par]
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 2);
      tc_fence_after();
      if (MODE == kPCQ) {
        if (leader) {
          issue_gemm<2, 128>(dcol, a_base, a_base + 32768, 16384, w_base, w_base + 32768, 16384, a.split, 0, FMT);
          mma_commit(d_full + 8 * (0 * 2 + par));
          mma_commit(a_free);   // the smem A image is dead: step 1 reads feat1 from tensor memory
        }
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 3);
        mbar_wait(a1_ready, a1_uses & 1);
        ++a1_uses;
        if (it >= 2) mbar_wait(d_free + 8 * par, (uint32_t)((u - 1) & 1));  // E1(t-2) drained D1[par]
        tc_fence_after();
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 4);
        if (leader) {
          issue_gemm_tchunk<4, 128>(dcol + 128, dcol, w_base + 65536, w_base + 65536 + 32768, 16384, a.split, FMT);
          mma_commit(d_full + 8 * (1 * 2 + par));
        }
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 5);
      } else if (MODE == kKV) {
        if (leader) {
          issue_gemm<2, 128>(dcol, a_base, a_base + 32768, 16384, w_base, w_base + 32768, 16384, a.split, 0, FMT);
          mma_commit(d_full + 8 * (0 * 2 + par));
          issue_gemm<2, 128>(dcol + 128, a_base, a_base + 32768, 16384, w_base + 65536, w_base + 65536 + 32768, 16384, a.split,
                             0, FMT);
          mma_commit(d_full + 8 * (1 * 2 + par));
          mma_commit(a_free);
        }
      } else {
        // Wm0: 64 x 128 (hi 16K | lo 16K, panel 8K)   Wm1: 64 x 64 (hi 8K | lo 8K)   Wm2: 128 x 64 (hi 16K | lo 16K)
        if (leader) {
          issue_gemm<2, 64>(dcol, a_base, a_base + 32768, 16384, w_base, w_base + 16384, 8192, a.split, 0, FMT);
          mma_commit(d_full + 8 * (0 * 2 + par));
          mma_commit(a_free);
        }
        mbar_wait(a1_ready, a1_uses & 1);
        ++a1_uses;
        tc_fence_after();
        if (leader) {
          issue_gemm_tchunk<2, 64>(dcol + 64, dcol, w_base + 32768, w_base + 32768 + 8192, 8192, a.split, FMT);
          mma_commit(d_full + 8 * (1 * 2 + par));
        }
        mbar_wait(a1_ready, a1_uses & 1);
        ++a1_uses;
        tc_fence_after();
        if (leader) {
          issue_gemm_tchunk<2, 128>(dcol + 128, dcol + 64, w_base + 49152, w_base + 49152 + 16384, 16384, a.split, FMT);
          mma_commit(d_full + 8 * (2 * 2 + par));
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // =================================== loaders: 7 warps, rows lw + 7 i ===================================
    const int lw = warp - 8;
    const bool stamp_ld = a.dbg != nullptr && blockIdx.x == 0 && tid == 256;
    int it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const long long row0 = tile * 128;
      float4 v[kChLoaderRows];
#pragma unroll
      for (int i = 0; i < kChLoaderRows; ++i) {
        const int rr = lw + kChLoaderWarps * i;
        const long long grow = row0 + rr;
        v[i] = (rr < 128 && grow < a.rows) ? __ldg(reinterpret_cast<const float4*>(a.in + grow * kC) + lane)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (stamp_ld) PDSC_STAMP1(a.dbg, it, 1, 0);
      if (it > 0) mbar_wait(a_free, (uint32_t)((it - 1) & 1));
      if (stamp_ld) PDSC_STAMP1(a.dbg, it, 1, 1);
#pragma unroll
      for (int i = 0; i < kChLoaderRows; ++i) {
        const int rr = lw + kChLoaderWarps * i;
        if (rr < 128) {
          uint32_t h0, l0, h1, l1;
          split_pair<FMT>(v[i].x, v[i].y, h0, l0);
          split_pair<FMT>(v[i].z, v[i].w, h1, l1);
          const uint32_t off = (uint32_t)(lane >> 4) * 16384u + sw128_offset((uint32_t)rr, (uint32_t)(lane & 15) * 4u);
          *reinterpret_cast<uint2*>(Abuf + off) = make_uint2(h0, h1);
          if (a.split) *reinterpret_cast<uint2*>(Abuf + 32768 + off) = make_uint2(l0, l1);
        }
      }
      if (stamp_ld) PDSC_STAMP1(a.dbg, it, 1, 2);
      fence_proxy_async_smem();
      mbar_arrive(a_ready);
      if (stamp_ld) PDSC_STAMP1(a.dbg, it, 1, 3);
      if (MODE == kMSG) {
        // residual tile: global -> smem without registers; 16-byte chunk c of row r lands at chunk (c & ~7) | ((c ^ r) & 7)
        if (it > 0) mbar_wait(r_free, (uint32_t)((it - 1) & 1));
#pragma unroll
        for (int i = 0; i < kChLoaderRows; ++i) {
          const int r = lw + kChLoaderWarps * i;
          if (r < 128) {
            const long long grow = row0 + r;
            const uint32_t dst = s0 + kChRes + (uint32_t)r * 512u + (uint32_t)(((lane & ~7) | ((lane ^ r) & 7)) << 4);
            const bool ok = grow < a.rows;
            cp_async16(dst, ok ? (const void*)(a.res + grow * kC + lane * 4) : (const void*)a.res, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(r_ready);
      }
    }
  } else {
    // =================================== epilogue: 2 warpgroups ===================================
    const int q4 = warp & 3, h = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
    uint8_t* stage = smem + kChStage + warp * 4096;          // PCQ / KV: [32 rows][128 B], 16-byte chunks XOR-swizzled by row
    uint8_t* resq = smem + kChRes + q4 * 32 * 512;           // MSG: the 32 residual rows of this lane quarter
    const int sub = lane >> 3, piece = lane & 7;             // read-out phase: rows sub + 4 i, 16-byte piece of the row
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    // one epilogue step of one tile (all per-tile addressing is derived here, so steps of different tiles can interleave)
    auto run_step = [&](const long long tile, const int it, const int step) {
      const int par = it & 1, u = it >> 1;
      if (stamp) PDSC_STAMP1(a.dbg, it, 3, 6);
      const uint32_t dcol = tmem + lane_base + (uint32_t)par * 256u;
      const long long row0 = tile * 128 + q4 * 32;          // first global row of this lane quarter
      // (set, index) of the quarter's first row: one division per tile, the 32 rows follow by comparison
      const int b0 = (int)((unsigned)row0 / (unsigned)a.N);
      const int n0 = (int)((unsigned)row0 - (unsigned)b0 * (unsigned)a.N);
      int my_b, my_n;
      locate_row(b0, n0, lane, a.N, my_b, my_n);
      // image destinations of the 8 rows this lane stores in the read-out phase (PCQ: Q, KV: K)
      uint8_t* img_row[8];
      uint32_t img_rx[8];
      if (MODE == kPCQ || MODE == kKV) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = sub + 4 * i;
          img_row[i] = nullptr;
          img_rx[i] = 0;
          if (row0 + rr < a.rows) {
            int bb, nn;
            locate_row(b0, n0, rr, a.N, bb, nn);
            if (MODE == kPCQ) {
              const uint32_t rit = (uint32_t)(nn & 127);
              img_row[i] = a.qimg + ((size_t)bb * a.QT + (nn >> 7)) * 65536 + (rit >> 3) * 1024u + (rit & 7u) * 128u;
              img_rx[i] = rit & 7u;
            } else {
              const uint32_t rit = (uint32_t)(nn & 63);
              img_row[i] = a.kvimg + ((size_t)bb * a.KT + (nn >> 6)) * 65536 + (rit >> 3) * 1024u + (rit & 7u) * 128u;
              img_rx[i] = rit & 7u;
            }
          }
        }
      }

      if (stamp) PDSC_STAMP1(a.dbg, it, 3, 7);
        if (stamp) PDSC_STAMP1(a.dbg, it, 2, step * 2);
        mbar_wait(d_full + 8 * (step * 2 + par), (uint32_t)(u & 1));
        if (stamp) PDSC_STAMP1(a.dbg, it, 2, step * 2 + 1);
        tc_fence_after();
        if (MODE == kMSG && step == 2) mbar_wait(r_ready, (uint32_t)(it & 1));
        const int ncols = (MODE == kMSG && step < 2) ? 64 : 128;
        const uint32_t dstep = (MODE == kMSG) ? (step == 0 ? 0u : (step == 1 ? 64u : 128u)) : (uint32_t)step * 128u;
        const float* bvec = bias + ((MODE == kMSG) ? (step == 0 ? 0 : (step == 1 ? 64 : 128)) : step * 128);
        const int cbeg = h * (ncols / 2), cend = cbeg + ncols / 2;
        const bool chained = (MODE == kPCQ && step == 0) || (MODE == kMSG && step < 2);   // feeds the next MMA
        if ((MODE == kPCQ && step == 1) || MODE == kKV) {
          // Q / K / V image -> HBM (V has the K format: rows = keys; the PV MMA reads it as an MN-major B operand).  This thread's 64 columns are exactly one 128-byte panel row (hi) and one (lo): stage
          // the warp's 32 rows x 128 B, then every store instruction writes four full 128-byte lines.
          if (stamp && step == 1) PDSC_STAMP1(a.dbg, it, 3, 0);
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t raw[32];
            tmem_ld32(dcol + dstep + cbeg + 32 * cc, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bvec + cbeg + 32 * cc + i);
              split_pair<FMT>(__uint_as_float(raw[i]) + bv.x, __uint_as_float(raw[i + 1]) + bv.y, hi[16 * cc + (i >> 1)],
                              lo[16 * cc + (i >> 1)]);
              split_pair<FMT>(__uint_as_float(raw[i + 2]) + bv.z, __uint_as_float(raw[i + 3]) + bv.w,
                              hi[16 * cc + (i >> 1) + 1], lo[16 * cc + (i >> 1) + 1]);
            }
          }
          if (stamp && step == 1) PDSC_STAMP1(a.dbg, it, 3, 1);
          const uint32_t panel_off = (uint32_t)h * ((MODE == kPCQ) ? 16384u : 8192u) + ((MODE == kKV && step == 1) ? 32768u : 0u);
          const uint32_t lo_off = (MODE == kPCQ) ? 32768u : 16384u;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            if (part == 1 && !a.split) break;
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t* src = part ? lo : hi;
              *reinterpret_cast<uint4*>(stage + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                  make_uint4(src[4 * g], src[4 * g + 1], src[4 * g + 2], src[4 * g + 3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = sub + 4 * i;
              const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * 128 + ((piece ^ (rr & 7)) << 4));
              if (img_row[i]) st_global_v4(img_row[i] + panel_off + (((uint32_t)piece ^ img_rx[i]) << 4) + (part ? lo_off : 0u), val);
            }
          }
          if (stamp && step == 1) PDSC_STAMP1(a.dbg, it, 3, 4);
        } else
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          const bool st0 = stamp && step == 0 && c0 == cbeg;   // timeline: loader slots 4-7 carry the step-0 epilogue detail
          if (st0) PDSC_STAMP1(a.dbg, it, 1, 4);
          uint32_t raw[32];
          tmem_ld32(dcol + dstep + c0, raw);
          tmem_ld_wait();
          if (st0) PDSC_STAMP1(a.dbg, it, 3, 2);
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bvec + c0 + i);
            x[i] = __uint_as_float(raw[i]) + bv.x;
            x[i + 1] = __uint_as_float(raw[i + 1]) + bv.y;
            x[i + 2] = __uint_as_float(raw[i + 2]) + bv.z;
            x[i + 3] = __uint_as_float(raw[i + 3]) + bv.w;
          }
          if (chained) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.f);
            // next step's A operand, in place over the columns just read: [hi 16 | lo 16]
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_pair<FMT>(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
            if (st0) PDSC_STAMP1(a.dbg, it, 3, 3);
            tmem_st16(dcol + dstep + c0, hi);
            if (a.split) tmem_st16(dcol + dstep + c0 + 16, lo);
            if (st0) PDSC_STAMP1(a.dbg, it, 1, 5);
          }

          if (MODE == kMSG && step == 2) {
            // feat = feat1 + fc_message(msg): add the residual in place in the smem tile (store-out below)
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int c = (c0 >> 2) + g;  // 16-byte chunk index within the 512-byte row
              float4* slot = reinterpret_cast<float4*>(resq + lane * 512 + (((c & ~7) | ((c ^ lane) & 7)) << 4));
              float4 rv = *slot;
              rv.x += x[g * 4]; rv.y += x[g * 4 + 1]; rv.z += x[g * 4 + 2]; rv.w += x[g * 4 + 3];
              *slot = rv;
            }
          }
          if (MODE == kPCQ && step == 0) {
            // feat1 fp32 -> HBM through the staging buffer (full-line stores)
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(stage + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                  make_float4(x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = sub + 4 * i;
              const long long g = row0 + rr;
              const float4 val = *reinterpret_cast<const float4*>(stage + rr * 128 + ((piece ^ (rr & 7)) << 4));
              if (g < a.rows) *reinterpret_cast<float4*>(a.out_f32 + g * kC + c0 + piece * 4) = val;
            }
            if (st0) PDSC_STAMP1(a.dbg, it, 1, 6);
          }
        }
        if (chained) {
          tmem_st_wait();        // A operand written through tcgen05.st, read by the tensor core
          tc_fence_before();
          mbar_arrive(a1_ready);
          if (stamp && step == 0) PDSC_STAMP1(a.dbg, it, 1, 7);
        }
    };
    if (MODE == kPCQ) {
      // Software-pipelined order  E0(t), E1(t-1), E0(t+1), E1(t), ...: the step-1 MMA of tile t (which needs E0(t)'s
      // output) runs under E1(t-1), so the epilogue warps never sit waiting for the tensor core.
      long long prev = -1;
      int it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        run_step(tile, it, 0);
        if (it > 0) {
          run_step(prev, it - 1, 1);
          tc_fence_before();
          mbar_arrive(d_free + 8 * ((it - 1) & 1));   // D1[par] drained
        }
        prev = tile;
      }
      if (it > 0) {
        run_step(prev, it - 1, 1);
        tc_fence_before();
        mbar_arrive(d_free + 8 * ((it - 1) & 1));
      }
    } else {
      int it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int par = it & 1;
        const long long row0 = tile * 128 + q4 * 32;
#pragma unroll
        for (int step = 0; step < kSteps; ++step) run_step(tile, it, step);
        if (MODE == kMSG) {
          // store the finished fp32 tile: one full 512-byte row per instruction, 16 rows per warp
          quarter_sync(q4);  // both column halves of these 32 rows are in place
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int rr = 16 * h + i;
            const long long g = row0 + rr;
            const float4 val = *reinterpret_cast<const float4*>(resq + rr * 512 + (((lane & ~7) | ((lane ^ rr) & 7)) << 4));
            if (g < a.rows) *reinterpret_cast<float4*>(a.out_f32 + g * kC + lane * 4) = val;
          }
          mbar_arrive(r_free);
        }
        tc_fence_before();
        mbar_arrive(d_free + 8 * par);  // D[par] drained
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 15) tmem_dealloc(tmem, 512);
}

}  // namespace pdsc
