// tc_chain: fused row-tile GEMM chains of the encoder's 1x1 convolutions on tcgen05 (included by encoder_tc.cu).
//
//   PCQ : feat  --W1,relu--> feat1 (fp32 -> HBM) --Wq--> Q image (HBM)
//   KV  : feat1 --Wk--> K image (HBM) ;  feat1 --Wv--> V image (HBM, same row-major format as K)
//   MSG : msg --Wm0,relu--> --Wm1,relu--> --Wm2--> + feat1 --> feat (fp32 -> HBM)
// (reference models/PointDSC.py:56-61 PointCN, :21-23/:36-38 projections, :12-20/:43-44 fc_message + residual)
//
// Persistent CTAs (one per SM), weights resident in shared memory, 128-row tiles.  An epilogue (TMEM -> registers -> bias /
// ReLU / hi-lo split -> stores) is a latency chain, so the two GEMM steps of a tile are drained by DIFFERENT warps working
// concurrently; with the stores off the shared-memory round trip the kernels run at 15-17 B/clk/SM of mixed read + write
// traffic (0.66-0.73 of the copy peak), and a version with two complete tile chains in flight was no faster.  Warp roles (512 threads, so that the
// register file divides into 128 registers per thread):
//   warps 0-3   group A: thread = row r (TMEM lane), all columns.  PCQ: PointCN step (feat1 -> HBM and, as 16-bit hi | lo
//               images, back into tensor memory for the Q GEMM).  KV: K image.  MSG: last step (+ residual, store).
//   warps 4-7   group B: thread = row r, all columns.  PCQ: Q image.  KV: V image.  MSG: the two 64-wide hidden steps.
//   warps 8-14  loaders : prefetch the NEXT tile's fp32 rows into registers (coalesced, rows lw + 7 i), convert them to
//               the swizzled 16-bit A image once the tensor core has released the buffer; MSG: also the residual tile
//   warp  15    MMA issuer (whole warp runs the control flow, one elected lane issues) + TMEM allocation
// Chained steps never go back through shared memory: the epilogue writes the next step's A operand over the accumulator
// columns it has just read, and the next MMA takes A FROM TENSOR MEMORY.  Chunk q (K elements 32 q .. 32 q + 31) of such
// an operand sits at columns 32 q .. 32 q + 31 of the producing accumulator as [hi: 16 columns | lo: 16 columns].
// Accumulators are double-buffered in TMEM by tile parity (2 x 256 columns).  The 16-bit operand images (Q, K, V) are stored
// straight from registers: a thread owns a row, a 32-column chunk is 64 bytes of its hi and 64 of its lo image row, written as
// 256-bit stores (STG.256, whole 32-byte sectors; the 128-byte swizzle only swaps the halves of a sector).  Same-box A/B against
// the staged form (transposition through shared memory, 8 rows x 64 B per instruction): chain kernels 3.51 -> 3.29 ms per step.
// feat1 (fp32, PCQ -> KV and MSG) lives in HBM in a BLOCKED layout keyed by the 128-row chain tile:
//     [tile][32-column chunk cc][128 rows][128 B], 16-byte piece q of row r at piece q ^ (r & 7)
// so that the 32 rows x 128 B a warp produces per chunk are 4 KB contiguous and leave as ONE bulk async copy (TMA engine) from
// the warp's shared-memory staging buffer — no read-back, no store instructions, and the warp does not wait for the data to
// leave the SM (two buffers per warp; the direct STG.256 form measured slower for these 128-byte pieces).  The KV loaders and
// the MSG residual loader read it with the same address function (blocked_f32_offset).
#pragma once
#include "tc_common.cuh"

namespace pdsc {

constexpr int kChainThreads = 512;
constexpr int kChLoaderWarps = 7, kChLoaderRows = 19;   // rows lw + 7 i, i < 19 (the last one only for lw < 2)
constexpr int kChA = 0;                          // A image: [hi p0 16K][hi p1 16K][lo p0 16K][lo p1 16K]
constexpr int kChW = 65536;                      // weight images (128 KB for PCQ / KV, 80 KB for MSG)
constexpr int kChStage = 65536 + 131072;         // PCQ: 4 warps x two 4 KB staging buffers of feat1 chunks (32 KB; unused by KV)
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

// the 64-byte half `m` (16 packed 16-bit pairs = 32 K elements) of this thread's 128-byte image row, whose 16-byte pieces are
// XOR-swizzled by x = (image row & 7): a 32-byte pair of pieces stays one 32-byte sector (its halves swap when x is odd)
__device__ __forceinline__ void store_image_half(uint8_t* row, uint32_t x, uint32_t m, const uint32_t (&v)[16]) {
  const uint32_t half = m ^ (x >> 2);
  const bool sw = (x & 1u) != 0u;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t pair = (uint32_t)k ^ ((x >> 1) & 1u);
    uint8_t* p = row + (half << 6) + (pair << 5);
    st_global_v8(p, sw ? v[8 * k + 4] : v[8 * k], sw ? v[8 * k + 5] : v[8 * k + 1], sw ? v[8 * k + 6] : v[8 * k + 2],
                 sw ? v[8 * k + 7] : v[8 * k + 3], sw ? v[8 * k] : v[8 * k + 4], sw ? v[8 * k + 1] : v[8 * k + 5],
                 sw ? v[8 * k + 2] : v[8 * k + 6], sw ? v[8 * k + 3] : v[8 * k + 7]);
  }
}

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

// byte offset of the 16-byte piece `piece` (0..31) of global row `g` in the blocked fp32 layout described in the header
__host__ __device__ __forceinline__ size_t blocked_f32_offset(long long g, uint32_t piece) {
  return (size_t)(g >> 7) * 65536 + (size_t)(piece >> 3) * 16384 + (size_t)(g & 127) * 128 + (size_t)(((piece & 7u) ^ ((uint32_t)g & 7u)) << 4);
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
  const uint32_t d_full = smem_u32(bars + 6);    // [step][parity] at + 8 * (step * 2 + parity)
  const uint32_t d0_free = smem_u32(bars + 12);  // [parity]  group A has drained its accumulator(s) of the tile
  const uint32_t d1_free = smem_u32(bars + 14);  // [parity]  group B has drained the second GEMM's accumulator
  // PCQ: group A's operand of the Q GEMM is in tensor memory, one barrier PER TILE PARITY.  The PointCN GEMM is issued one
  // tile ahead, so a fast warp of group A can finish tile t + 1 while a slow one is still on tile t: on a single barrier
  // its second arrival would be counted towards tile t's phase, the phase would complete without the slow warp, and the
  // Q GEMM would read that warp's 32 rows before they were written (the round-1 "one corrupted lane quarter about once in
  // four forwards": profiles/r02_determinism_campaign.txt).  With two barriers a warp's next arrival on the same barrier
  // is for tile t + 2, whose PointCN GEMM is only issued after the MMA warp has passed tile t's wait.
  const uint32_t a1_par = smem_u32(bars + 16);   // [parity]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_base = s0 + kChA, w_base = s0 + kChW;
  // this mode's biases, packed: PCQ b1|bq, KV bk|bv, MSG bm0|bm1|bm2
  constexpr int kBiasSrc = (MODE == kPCQ) ? kB1 : (MODE == kKV) ? kBk : kBm0;
  const int N = a.N;
  const long long rows = a.rows;

  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(bar_w, 1);
    mbar_init(a_ready, kChLoaderWarps * 32);
    mbar_init(a1_ready, 128);
    mbar_init(a_free, 1);
    mbar_init(r_ready, kChLoaderWarps * 32);
    mbar_init(r_free, 128);
    for (int i = 0; i < 6; ++i) mbar_init(d_full + 8 * i, 1);
    mbar_init(d0_free, 128); mbar_init(d0_free + 8, 128);
    mbar_init(d1_free, 128); mbar_init(d1_free + 8, 128);
    mbar_init(a1_par, 128); mbar_init(a1_par + 8, 128);
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
  const long long num_tiles = (rows + 127) / 128;

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
      if (MODE != kPCQ) mbar_wait(a_ready, (uint32_t)(it & 1));
      if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 1);
      if (MODE == kPCQ) {
        // Issue order  M0(t), M0(t+1)?, ... is software-pipelined: the PointCN GEMM of tile t + 1 goes out BEFORE the wait for
        // group A's operand of tile t, so group A finds its next accumulator ready the moment it finishes a tile.
        // D0[par] is free: the a1_ready wait of tile t - 2 covered group A's drain, and issue_m0 waits for the completion of
        // the Q GEMM of tile t - 2 (the last reader of the operand parked there).
        auto issue_m0 = [&](int itx) {
          mbar_wait(a_ready, (uint32_t)(itx & 1));
          // M0(itx) overwrites D0[itx & 1], which the Q GEMM of tile itx - 2 reads as its A operand, and that MMA directly
          // precedes this one in the issue order: wait for its COMPLETION (its own commit barrier; the next phase of that
          // barrier is the Q GEMM of tile itx, which this warp has not issued yet) — see the rule in tc_common.cuh.
          if (itx >= 2) mbar_wait(d_full + 8 * (1 * 2 + (itx & 1)), (uint32_t)(((itx >> 1) - 1) & 1));
          tc_fence_after();
          if (leader) {
            const uint32_t dc = tmem + (uint32_t)(itx & 1) * 256u;
            issue_gemm<2, 128>(dc, a_base, a_base + 32768, 16384, w_base, w_base + 32768, 16384, a.split, 0, FMT);
            mma_commit(d_full + 8 * (0 * 2 + (itx & 1)));
            mma_commit(a_free);   // the smem A image is dead: step 1 reads feat1 from tensor memory
          }
        };
        if (it == 0) issue_m0(0);
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 3);
        // M0(t+1) needs D0[par ^ 1]: drained by group A in tile t - 1 (covered by that tile's a1_ready wait) and no longer
        // read by the Q GEMM of tile t - 1 (completion awaited inside issue_m0).
        if (tile + gridDim.x < num_tiles) issue_m0(it + 1);
        mbar_wait(a1_par + 8 * par, (uint32_t)(u & 1));
        if (it >= 2) mbar_wait(d1_free + 8 * par, (uint32_t)((u - 1) & 1));  // group B drained D1[par] of tile t - 2
        tc_fence_after();
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 4);
        if (leader) {
          issue_gemm_tchunk<4, 128>(dcol + 128, dcol, w_base + 65536, w_base + 65536 + 32768, 16384, a.split, FMT);
          mma_commit(d_full + 8 * (1 * 2 + par));
        }
        if (stamp_mma) PDSC_STAMP1(a.dbg, it, 0, 5);
      } else if (MODE == kKV) {
        if (it >= 2) {
          mbar_wait(d0_free + 8 * par, (uint32_t)((u - 1) & 1));
          mbar_wait(d1_free + 8 * par, (uint32_t)((u - 1) & 1));
        }
        tc_fence_after();
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
        // D0 / D1 [par] are free: the a1_ready waits of tile t - 2 covered group B's drains, and the MMAs of that tile that read
        // them as A operands have COMPLETED (group B arrived on a1_ready of tile t - 1 only after it saw that tile's first
        // commit, which tracks every MMA issued before it)
        tc_fence_after();
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
        if (it >= 2) mbar_wait(d0_free + 8 * par, (uint32_t)((u - 1) & 1));  // group A drained D2[par] of tile t - 2
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
      // blocked feat1 (KV: the input; MSG: the residual): this tile's 64 KB block, this lane's 32-column chunk
      const uint8_t* blk_base = reinterpret_cast<const uint8_t*>(MODE == kKV ? a.in : a.res) + (size_t)tile * 65536 + (size_t)(lane >> 3) * 16384;
      float4 v[kChLoaderRows];
#pragma unroll
      for (int i = 0; i < kChLoaderRows; ++i) {
        const int rr = lw + kChLoaderWarps * i;
        const long long grow = row0 + rr;
        if (MODE == kKV)   // feat1: blocked layout (tile base + a 32-bit offset: row rr of the tile, this lane's 16-byte piece)
          v[i] = (rr < 128 && grow < rows)
                     ? __ldg(reinterpret_cast<const float4*>(blk_base + (uint32_t)rr * 128u + ((((uint32_t)lane & 7u) ^ ((uint32_t)rr & 7u)) << 4)))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        else
          v[i] = (rr < 128 && grow < rows) ? __ldg(reinterpret_cast<const float4*>(a.in + grow * kC) + lane)
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
            const bool ok = grow < rows;
            cp_async16(dst, ok ? (const void*)(blk_base + (uint32_t)r * 128u + ((((uint32_t)lane & 7u) ^ ((uint32_t)r & 7u)) << 4))
                                 : (const void*)a.res, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(r_ready);
      }
    }
  } else if (warp >= 4) {
    // =================================== group B: thread = row ===================================
    if (MODE == kPCQ || MODE == kKV) {
      const int q4 = warp & 3;
      const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
      const float* bvec = bias + 128;
      uint8_t* const img = (MODE == kPCQ) ? a.qimg : a.kvimg;
      const uint32_t panel_bytes = (MODE == kPCQ) ? 16384u : 8192u;
      const uint32_t lo_off = (MODE == kPCQ) ? 32768u : 16384u;
      const uint32_t base_off = (MODE == kPCQ) ? 0u : 32768u;   // V sits behind K in the tile's image
      int it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int par = it & 1, u = it >> 1;
        const uint32_t dcol = tmem + lane_base + (uint32_t)par * 256u + 128u;
        const long long row0 = tile * 128 + q4 * 32;
        const int b0 = (int)((unsigned)row0 / (unsigned)N);
        const int n0 = (int)((unsigned)row0 - (unsigned)b0 * (unsigned)N);
        // this thread's row of the image (thread = row: it stores its own 64-byte pieces as two 32-byte sectors each)
        size_t own_off;
        uint32_t own_x;
        const bool own_ok = row0 + lane < rows;
        {
          int bb, nn;
          locate_row(b0, n0, lane, N, bb, nn);
          const uint32_t rit = (MODE == kPCQ) ? (uint32_t)(nn & 127) : (uint32_t)(nn & 63);
          own_off = ((MODE == kPCQ) ? ((size_t)bb * a.QT + (nn >> 7)) : ((size_t)bb * a.KT + (nn >> 6))) * 65536 + rit * 128u;
          own_x = rit & 7u;
        }
        mbar_wait(d_full + 8 * (1 * 2 + par), (uint32_t)(u & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {           // 32 columns = half of one 128-byte image row of panel c >> 1
          uint32_t raw[32];
          tmem_ld32(dcol + 32 * c, raw);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bvec + 32 * c + i);
            split_pair<FMT>(__uint_as_float(raw[i]) + bv.x, __uint_as_float(raw[i + 1]) + bv.y, hi[i >> 1], lo[i >> 1]);
            split_pair<FMT>(__uint_as_float(raw[i + 2]) + bv.z, __uint_as_float(raw[i + 3]) + bv.w, hi[(i >> 1) + 1], lo[(i >> 1) + 1]);
          }
          if (c == 3) {
            tc_fence_before();
            mbar_arrive(d1_free + 8 * par);      // D1[par] drained (the values live in registers now)
          }
          const uint32_t poff = base_off + (uint32_t)(c >> 1) * panel_bytes;
          const uint32_t m = (uint32_t)(c & 1);  // which 64-byte half of the 128-byte row
          if (own_ok) {
            store_image_half(img + own_off + poff, own_x, m, hi);
            if (a.split) store_image_half(img + own_off + poff + lo_off, own_x, m, lo);
          }
        }
      }
    }
    if (MODE == kMSG) {
      // ---- the two hidden steps of fc_message: 64 columns each, result parked in place as the next step's A operand ----
      const int q4 = warp & 3;
      const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
      int it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int par = it & 1, u = it >> 1;
        const uint32_t dcol = tmem + lane_base + (uint32_t)par * 256u;
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
          mbar_wait(d_full + 8 * (step * 2 + par), (uint32_t)(u & 1));
          tc_fence_after();
          const uint32_t dstep = step == 0 ? 0u : 64u;
          const float* bvec = bias + (step == 0 ? 0 : 64);
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t raw[32];
            tmem_ld32(dcol + dstep + c0, raw);
            tmem_ld_wait();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bvec + c0 + i);
              split_pair<FMT>(fmaxf(__uint_as_float(raw[i]) + bv.x, 0.f), fmaxf(__uint_as_float(raw[i + 1]) + bv.y, 0.f), hi[i >> 1], lo[i >> 1]);
              split_pair<FMT>(fmaxf(__uint_as_float(raw[i + 2]) + bv.z, 0.f), fmaxf(__uint_as_float(raw[i + 3]) + bv.w, 0.f), hi[(i >> 1) + 1],
                              lo[(i >> 1) + 1]);
            }
            tmem_st16(dcol + dstep + c0, hi);
            if (a.split) tmem_st16(dcol + dstep + c0 + 16, lo);
          }
          tmem_st_wait();        // A operand written through tcgen05.st, read by the tensor core
          tc_fence_before();
          mbar_arrive(a1_ready);
        }
      }
    }
  } else {
    // =================================== group A: thread = row ===================================
    const int q4 = warp & 3;
    const uint32_t lane_base = ((uint32_t)(q4 * 32)) << 16;
    uint8_t* stage = smem + kChStage + warp * 8192;          // PCQ: two 4 KB buffers per warp (feat1 chunks on their way to the TMA engine)
    uint8_t* resq = smem + kChRes + q4 * 32 * 512;           // MSG: the 32 residual rows of this lane quarter
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int par = it & 1, u = it >> 1;
      const uint32_t dcol = tmem + lane_base + (uint32_t)par * 256u;
      const long long row0 = tile * 128 + q4 * 32;          // first global row of this lane quarter
      if (MODE == kPCQ) {
        // ---- PointCN: feat1 = relu(D0 + b1)  ->  HBM (fp32) and tensor memory (16-bit hi | lo operand of the Q GEMM) ----
        if (stamp) PDSC_STAMP1(a.dbg, it, 2, 0);
        mbar_wait(d_full + 8 * (0 * 2 + par), (uint32_t)(u & 1));
        if (stamp) PDSC_STAMP1(a.dbg, it, 2, 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = 32 * cc;
          uint32_t raw[32];
          tmem_ld32(dcol + c0, raw);
          tmem_ld_wait();
          uint32_t xb[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias + c0 + i);
            xb[i] = __float_as_uint(fmaxf(__uint_as_float(raw[i]) + bv.x, 0.f));
            xb[i + 1] = __float_as_uint(fmaxf(__uint_as_float(raw[i + 1]) + bv.y, 0.f));
            xb[i + 2] = __float_as_uint(fmaxf(__uint_as_float(raw[i + 2]) + bv.z, 0.f));
            xb[i + 3] = __float_as_uint(fmaxf(__uint_as_float(raw[i + 3]) + bv.w, 0.f));
          }
          {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_pair<FMT>(__uint_as_float(xb[2 * i]), __uint_as_float(xb[2 * i + 1]), hi[i], lo[i]);
            tmem_st16(dcol + c0, hi);
            if (a.split) tmem_st16(dcol + c0 + 16, lo);
          }
          if (cc == 3) {
            tmem_st_wait();        // A operand written through tcgen05.st, read by the tensor core
            tc_fence_before();
            mbar_arrive(a1_par + 8 * par);
            if (stamp) PDSC_STAMP1(a.dbg, it, 2, 2);
          }
          // fp32 rows -> HBM (blocked layout): park the chunk's 32 rows x 128 B in one of the warp's two staging buffers and hand
          // them to the TMA engine as one contiguous 4 KB copy; the buffer used two chunks ago must have been read by then
          {
            uint8_t* buf = stage + (cc & 1) * 4096;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(xb[4 * q], xb[4 * q + 1], xb[4 * q + 2], xb[4 * q + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            const long long valid = rows - row0;      // rows of this lane quarter that exist (the last tile may be ragged)
            if (lane == 0 && valid > 0) {
              bulk_s2g(reinterpret_cast<uint8_t*>(a.out_f32) + (size_t)tile * 65536 + (size_t)cc * 16384 + (size_t)q4 * 4096, smem_u32(buf),
                       (uint32_t)(valid < 32 ? valid : 32) * 128u);
              bulk_commit();
            }
          }
        }
        if (stamp) PDSC_STAMP1(a.dbg, it, 2, 3);
      } else if (MODE == kKV) {
        // ---- K image: 32 columns = half of one 128-byte image row of panel c >> 1 ----
        const int b0 = (int)((unsigned)row0 / (unsigned)N);
        const int n0 = (int)((unsigned)row0 - (unsigned)b0 * (unsigned)N);
        size_t own_off;
        uint32_t own_x;
        const bool own_ok = row0 + lane < rows;
        {
          int bb, nn;
          locate_row(b0, n0, lane, N, bb, nn);
          const uint32_t rit = (uint32_t)(nn & 63);
          own_off = ((size_t)bb * a.KT + (nn >> 6)) * 65536 + rit * 128u;
          own_x = rit & 7u;
        }
        mbar_wait(d_full + 8 * (0 * 2 + par), (uint32_t)(u & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t raw[32];
          tmem_ld32(dcol + 32 * c, raw);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias + 32 * c + i);
            split_pair<FMT>(__uint_as_float(raw[i]) + bv.x, __uint_as_float(raw[i + 1]) + bv.y, hi[i >> 1], lo[i >> 1]);
            split_pair<FMT>(__uint_as_float(raw[i + 2]) + bv.z, __uint_as_float(raw[i + 3]) + bv.w, hi[(i >> 1) + 1], lo[(i >> 1) + 1]);
          }
          if (c == 3) {
            tc_fence_before();
            mbar_arrive(d0_free + 8 * par);
          }
          const uint32_t poff = (uint32_t)(c >> 1) * 8192u;
          const uint32_t m = (uint32_t)(c & 1);
          if (own_ok) {
            store_image_half(a.kvimg + own_off + poff, own_x, m, hi);
            if (a.split) store_image_half(a.kvimg + own_off + poff + 16384u, own_x, m, lo);
          }
        }
      } else {
        // ---- MSG, last step: feat = feat1 + (D2 + bm2), through the residual tile in shared memory ----
        mbar_wait(d_full + 8 * (2 * 2 + par), (uint32_t)(u & 1));
        tc_fence_after();
        mbar_wait(r_ready, (uint32_t)(it & 1));
        const float* bvec = bias + 128;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(dcol + 128 + c0, raw);
          tmem_ld_wait();
          if (c0 == 96) {
            tc_fence_before();
            mbar_arrive(d0_free + 8 * par);  // D2[par] drained
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bv = *reinterpret_cast<const float4*>(bvec + c0 + 4 * g);
            const int c = (c0 >> 2) + g;  // 16-byte chunk index within the 512-byte row
            float4* slot = reinterpret_cast<float4*>(resq + lane * 512 + (((c & ~7) | ((c ^ lane) & 7)) << 4));
            float4 rv = *slot;
            rv.x += __uint_as_float(raw[g * 4]) + bv.x; rv.y += __uint_as_float(raw[g * 4 + 1]) + bv.y;
            rv.z += __uint_as_float(raw[g * 4 + 2]) + bv.z; rv.w += __uint_as_float(raw[g * 4 + 3]) + bv.w;
            *slot = rv;
          }
        }
        // store the finished fp32 rows: one full 512-byte row per instruction
        __syncwarp();
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const long long g = row0 + rr;
          const float4 val = *reinterpret_cast<const float4*>(resq + rr * 512 + (((lane & ~7) | ((lane ^ rr) & 7)) << 4));
          if (g < rows) *reinterpret_cast<float4*>(a.out_f32 + g * kC + lane * 4) = val;
        }
        mbar_arrive(r_free);
      }
    }
  }
  if (MODE == kPCQ && warp < 4 && lane == 0) bulk_wait_all<0>();   // the staged feat1 chunks have left shared memory (and are complete)
  tc_fence_before();
  __syncthreads();
  if (warp == 15) tmem_dealloc(tmem, 512);
}

}  // namespace pdsc
