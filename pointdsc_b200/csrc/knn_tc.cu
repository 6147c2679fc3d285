// a7 on the tensor cores: feature-space distances of the SEED rows to all N correspondences of their set.
//
// Reference: models/common.py:58-61 (`2 - 2 * x @ x^T` over L2-normalised features) restricted to the rows gathered at
// models/PointDSC.py:252.  D[s][j] = 2 - 2 <f_seed(s), f_j>  is a [S x N x 128] GEMM per set; it runs as fp16 hi/lo split
// products (22 significant bits per operand, fp32 accumulation in TMEM — the same fp32-grade arithmetic as the encoder's
// default mode) and replaces the gather + SIMT SGEMM of the exact-arithmetic path (26 MFLOP per set).
//
// One CTA = up to 128 seed rows of one set x all keys, 64-key tiles.  Warp roles (384 threads):
//   warps 0-3  owners : thread = seed row = TMEM lane.  Load the seed's feature row, split it and write it to TENSOR
//              MEMORY as the A operand (no shared memory for A); per tile read D, form 2 - 2 acc, store the row segment
//   warps 4-7  loaders: prefetch the next 64 key rows (fp32) into registers, convert to the swizzled K-major B image
//   warp  8    MMA issuer (elect.sync lane), D double-buffered in TMEM
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace pdsc {

constexpr int kKnnThreads = 384;
constexpr int kKnnB = 0;                       // 2 stages x 32 KB: [hi p0 8K][hi p1 8K][lo p0 8K][lo p1 8K]
constexpr int kKnnBars = 65536;
constexpr int kKnnSmem = kKnnBars + 256;

__global__ void __launch_bounds__(kKnnThreads, 1) knn_dist_tc_kernel(const float* __restrict__ normed,
                                                                     const int32_t* __restrict__ seeds,
                                                                     float* __restrict__ dist, int N, int S, int tiles_per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kKnnBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t a_ready = smem_u32(bars + 0);
  const uint32_t b_ready = smem_u32(bars + 1), b_free = smem_u32(bars + 3);   // [2]
  const uint32_t d_full = smem_u32(bars + 5), d_free = smem_u32(bars + 7);    // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, s_base = blockIdx.x * 128;
  // key tiles [t0, t0 + T) of this CTA: blockIdx.z splits the keys when seed-row tiles x sets alone would leave SMs idle
  const int t0 = blockIdx.z * tiles_per_cta;
  const int T = min(tiles_per_cta, (N + 63) / 64 - t0);
  constexpr int FMT = kFmtF16;

  if (tid == 0) {
    if (s0 & 1023u) {
      printf("pointdsc_b200: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(a_ready, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_ready + 8 * i, 128); mbar_init(b_free + 8 * i, 1);
      mbar_init(d_full + 8 * i, 1); mbar_init(d_free + 8 * i, 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tA = tmem + 128;   // D buffers at +0 / +64; A hi image at +128 (64 columns), lo image at +192

  if (warp == 8) {
    // =================================== MMA issuer ===================================
    const bool leader = elect_one();
    mbar_wait(a_ready, 0);
    for (int t = 0; t < T; ++t) {
      const int st = t & 1, use = t >> 1;
      mbar_wait(b_ready + 8 * st, (uint32_t)(use & 1));
      if (use > 0) mbar_wait(d_free + 8 * st, (uint32_t)((use - 1) & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t bb = s0 + kKnnB + st * 32768;
        const uint32_t idesc = idesc_f16kind(128, 64, FMT);
        constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);
        uint32_t acc = 0;
#pragma unroll
        for (int p3 = 0; p3 < 3; ++p3) {          // hi*hi, hi*lo, lo*hi
          const uint32_t at = tA + (p3 == 2 ? 64 : 0);
          const uint32_t bo = bb + (p3 == 1 ? 16384 : 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t blo = (((bo + (ks >> 2) * 8192 + (ks & 3) * 32) >> 4) & 0x3FFFu) | (1u << 16);
            asm volatile(
                "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
                "mov.b64 db, {%2, %5};\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
                ::"r"(tmem + 64 * st), "r"(at + ks * 8), "r"(blo), "r"(idesc), "r"(acc), "r"(kDescHi)
                : "memory");
            acc = 1;
          }
        }
        mma_commit(d_full + 8 * st);
        mma_commit(b_free + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // =================================== loaders: 4 warps x 16 key rows ===================================
    const int lw = warp - 4;
    const float* rows = normed + (size_t)b * N * kC;
    for (int t = 0; t < T; ++t) {
      const int st = t & 1, use = t >> 1;
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int key = (t0 + t) * 64 + lw * 16 + i;
        v[i] = (key < N) ? __ldg(reinterpret_cast<const float4*>(rows + (size_t)key * kC) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (use > 0) mbar_wait(b_free + 8 * st, (uint32_t)((use - 1) & 1));
      uint8_t* Bs = smem + kKnnB + st * 32768;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t h0, l0, h1, l1;
        split_pair<FMT>(v[i].x, v[i].y, h0, l0);
        split_pair<FMT>(v[i].z, v[i].w, h1, l1);
        const uint32_t off = (uint32_t)(lane >> 4) * 8192u + sw128_offset((uint32_t)(lw * 16 + i), (uint32_t)(lane & 15) * 4u);
        *reinterpret_cast<uint2*>(Bs + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Bs + 16384 + off) = make_uint2(l0, l1);
      }
      fence_proxy_async_smem();
      mbar_arrive(b_ready + 8 * st);
    }
  } else if (warp < 4) {
    // =================================== owners ===================================
    const int r = tid;                         // TMEM lane
    const int s = s_base + r;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const bool live = s < S;
    {
      const float* frow = normed;
      if (live) {
        int idx = seeds[(size_t)b * S + s];
        idx = min(max(idx, 0), N - 1);
        frow = normed + ((size_t)b * N + idx) * kC;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {            // 32 channels at a time
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) f = __ldg(reinterpret_cast<const float4*>(frow + c * 32) + q);
          split_pair<FMT>(f.x, f.y, hi[2 * q], lo[2 * q]);
          split_pair<FMT>(f.z, f.w, hi[2 * q + 1], lo[2 * q + 1]);
        }
        tmem_st16(tA + lane_base + c * 16, hi);
        tmem_st16(tA + lane_base + 64 + c * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(a_ready);
    }
    float* drow = dist + ((size_t)b * S + (live ? s : 0)) * N;
    const bool vec_ok = (N & 3) == 0;
    for (int t = 0; t < T; ++t) {
      const int st = t & 1, use = t >> 1;
      mbar_wait(d_full + 8 * st, (uint32_t)(use & 1));
      tc_fence_after();
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        uint32_t raw[32];
        tmem_ld32(tmem + lane_base + 64 * st + 32 * hcol, raw);
        tmem_ld_wait();
        const int j0 = (t0 + t) * 64 + 32 * hcol;
        if (live) {
          if (vec_ok && j0 + 32 <= N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(drow + j0 + 4 * q) =
                  make_float4(fmaf(-2.f, __uint_as_float(raw[4 * q]), 2.f), fmaf(-2.f, __uint_as_float(raw[4 * q + 1]), 2.f),
                              fmaf(-2.f, __uint_as_float(raw[4 * q + 2]), 2.f), fmaf(-2.f, __uint_as_float(raw[4 * q + 3]), 2.f));
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (j0 + q < N) drow[j0 + q] = fmaf(-2.f, __uint_as_float(raw[q]), 2.f);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(d_free + 8 * st);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

void launch_knn_dist_tc(const float* normed, const int32_t* seeds, float* dist, int B, int N, int S, cudaStream_t st) {
  if (S <= 0) return;
  ensure_dynamic_smem(reinterpret_cast<const void*>(knn_dist_tc_kernel), kKnnSmem);
  const int T = (N + 63) / 64, ctas = ((S + 127) / 128) * B, sms = device_sm_count();
  int chunks = ctas >= sms ? 1 : (sms + ctas - 1) / ctas;
  if (chunks > T) chunks = T;
  const int per = (T + chunks - 1) / chunks;
  chunks = (T + per - 1) / per;
  knn_dist_tc_kernel<<<dim3((S + 127) / 128, B, chunks), kKnnThreads, kKnnSmem, st>>>(normed, seeds, dist, N, S, per);
}

}  // namespace pdsc
