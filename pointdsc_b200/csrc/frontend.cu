// f1 — the correspondence front end on the device.
//
// Reference: datasets/ThreeDMatch.py:283-291 (matching, optional mutual check) and :299-308 (network input); the same lines
// in datasets/KITTI.py:80-114 and demo_registration.py:101-108 (no mutual check there):
//     distance   = sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6)            in the descriptors' dtype (fp32 FCGF, fp64 FPFH)
//     source_idx = argmin(distance, axis=1)                                 numpy: the FIRST minimum
//     mutual     : keep i iff argmin(distance, axis=0)[source_idx[i]] == i
//     corr_pos   = concat(src_keypts[corr[:,0]], tgt_keypts[corr[:,1]]) - mean over the kept correspondences
//
// Kernels (one pair per call, Ns source and Nt target key points, descriptor dimension D <= 64):
//   match_rows_kernel<T>   thread = source row (descriptor in registers-by-smem), target descriptors staged through shared
//                          memory in tiles and read as broadcasts; the row keeps a running (distance, index) with a strict '<',
//                          which is numpy's first-minimum rule.  The square root is applied BEFORE the comparison, as the
//                          reference does (it merges distances that differ by less than an ulp of the root).  With the mutual
//                          check, column minima are merged across CTAs with atomicMin on the distance's bit pattern
//                          (distances are >= 0, so the IEEE bit patterns order like the values).
//   match_cols_kernel<T>   second pass for the mutual check: the lowest source index among those that attain a column's minimum
//                          (every distance is recomputed by the same instruction sequence, so equality is exact).
//   compact_center_kernel  one CTA: ordered compaction of the kept pairs (ascending source index), gather of the key points,
//                          column means in fp64, centred corr_pos — written in the layout pdsc_forward consumes.
// The dot product is accumulated in ascending channel order with one FMA per channel in T; the reference's BLAS fixes no
// accumulation order, so indices are only comparable where the best and the second-best distance are separated (tests).
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace pdsc {

constexpr int kMatchRows = 128;   // source rows per CTA (one per thread)
constexpr int kMatchTile = 64;    // target rows per shared-memory tile
constexpr int kMatchMaxD = 64;

template <typename T> struct BitsOf;
template <> struct BitsOf<float> {
  using type = unsigned int;
  __device__ static unsigned int get(float x) { return __float_as_uint(x); }
};
template <> struct BitsOf<double> {
  using type = unsigned long long;
  __device__ static unsigned long long get(double x) { return (unsigned long long)__double_as_longlong(x); }
};

template <typename T>
__device__ __forceinline__ T feature_distance(T dot) {
  // np.sqrt(2 - 2 * dot + 1e-6): python scalars adopt the array dtype, every operation rounded separately
  if (sizeof(T) == 4) return (T)__fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, (float)dot)), 1e-6f));
  return (T)__dsqrt_rn(__dadd_rn(__dsub_rn(2.0, __dmul_rn(2.0, (double)dot)), 1e-6));
}

// PASS 0: row minima (+ column minimum VALUES when col_min != nullptr).   PASS 1: column argmin given the column minima.
template <typename T, int PASS>
__global__ void __launch_bounds__(kMatchRows) match_kernel(const T* __restrict__ src_desc, const T* __restrict__ tgt_desc, int Ns,
                                                           int Nt, int D, int32_t* __restrict__ row_idx,
                                                           typename BitsOf<T>::type* __restrict__ col_min,
                                                           int32_t* __restrict__ col_idx) {
  extern __shared__ __align__(16) unsigned char match_smem[];
  T* s_src = reinterpret_cast<T*>(match_smem);       // [D][kMatchRows]   (channel-major: conflict-free per-thread reads)
  T* s_tgt = s_src + (size_t)D * kMatchRows;          // [kMatchTile][D]   (read as broadcasts)
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kMatchRows + tid;
  for (int e = tid; e < D * kMatchRows; e += kMatchRows) {
    const int r = e / D, d = e % D;                   // coalesced over the row-major descriptor block
    const int gi = blockIdx.x * kMatchRows + r;
    s_src[(size_t)d * kMatchRows + r] = gi < Ns ? src_desc[(size_t)gi * D + d] : (T)0;
  }
  T best = (T)0;
  int best_j = -1;
  for (int j0 = 0; j0 < Nt; j0 += kMatchTile) {
    __syncthreads();
    const int cnt = min(kMatchTile, Nt - j0);
    for (int e = tid; e < cnt * D; e += kMatchRows) s_tgt[e] = tgt_desc[(size_t)j0 * D + e];
    __syncthreads();
    if (i < Ns) {
      for (int t = 0; t < cnt; ++t) {
        T acc = (T)0;
        const T* tg = s_tgt + (size_t)t * D;
        for (int d = 0; d < D; ++d) {
          if (sizeof(T) == 4) acc = (T)__fmaf_rn((float)s_src[(size_t)d * kMatchRows + tid], (float)tg[d], (float)acc);
          else acc = (T)__fma_rn((double)s_src[(size_t)d * kMatchRows + tid], (double)tg[d], (double)acc);
        }
        const T dist = feature_distance<T>(acc);
        const int j = j0 + t;
        if (PASS == 0) {
          if (best_j < 0 || dist < best) { best = dist; best_j = j; }    // strict '<': the first minimum wins
          if (col_min) atomicMin(col_min + j, BitsOf<T>::get(dist));
        } else {
          if (BitsOf<T>::get(dist) == col_min[j]) atomicMin(col_idx + j, i);
        }
      }
    }
  }
  if (PASS == 0 && i < Ns) row_idx[i] = best_j;
}

template <typename Bits>
__global__ void match_init_kernel(Bits* col_min, int32_t* col_idx, int Nt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Nt) {
    col_min[j] = ~(Bits)0;
    col_idx[j] = 0x7FFFFFFF;
  }
}

// One CTA.  keep[i] = !mutual || col_idx[row_idx[i]] == i ; ordered compaction ; gather ; centre.
constexpr int kCompactThreads = 1024;
__global__ void __launch_bounds__(kCompactThreads) compact_center_kernel(const int32_t* __restrict__ row_idx,
                                                                         const int32_t* __restrict__ col_idx,
                                                                         const float* __restrict__ src_keypts,
                                                                         const float* __restrict__ tgt_keypts, int Ns, int mutual,
                                                                         int32_t* __restrict__ corr, int32_t* __restrict__ count,
                                                                         float* __restrict__ corr_pos, float* __restrict__ out_src,
                                                                         float* __restrict__ out_tgt) {
  __shared__ int warp_tot[kCompactThreads / 32];
  __shared__ int base_s;
  __shared__ double sums[6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  if (tid < 6) sums[tid] = 0.0;
  __syncthreads();
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i0 = 0; i0 < Ns; i0 += kCompactThreads) {
    const int i = i0 + tid;
    int j = -1;
    bool keep = false;
    if (i < Ns) {
      j = row_idx[i];
      keep = j >= 0 && (!mutual || col_idx[j] == i);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (keep) {
      const int m = before + in_warp;
      corr[2 * m] = i;
      corr[2 * m + 1] = j;
      float v[6];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[c] = src_keypts[(size_t)i * 3 + c];
        v[3 + c] = tgt_keypts[(size_t)j * 3 + c];
        out_src[(size_t)m * 3 + c] = v[c];
        out_tgt[(size_t)m * 3 + c] = v[3 + c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        corr_pos[(size_t)m * 6 + c] = v[c];
        acc[c] += (double)v[c];
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < kCompactThreads / 32; ++w) tot += warp_tot[w];
      base_s += tot;
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double s = warp_sum(acc[c]);
    if (lane == 0) atomicAdd(&sums[c], s);
  }
  __syncthreads();
  const int M = base_s;
  if (tid == 0) *count = M;
  // corr_pos - corr_pos.mean(0)   (ThreeDMatch.py:305-308): mean rounded to fp32, one rounded subtraction per element
  for (int e = tid; e < M * 6; e += kCompactThreads) {
    const float mean = (float)(sums[e % 6] / (double)M);
    corr_pos[e] = __fsub_rn(corr_pos[e], mean);
  }
}

template <typename T>
static void launch_match_t(const T* src_desc, const T* tgt_desc, const float* src_keypts, const float* tgt_keypts, int Ns, int Nt,
                           int D, int mutual, void* scratch, int32_t* corr, int32_t* count, float* corr_pos, float* out_src,
                           float* out_tgt, cudaStream_t st) {
  using Bits = typename BitsOf<T>::type;
  // scratch: row_idx [Ns] int32 | col_idx [Nt] int32 | col_min [Nt] Bits (8-byte aligned)
  int32_t* row_idx = static_cast<int32_t*>(scratch);
  int32_t* col_idx = row_idx + Ns;
  Bits* col_min = reinterpret_cast<Bits*>((reinterpret_cast<uintptr_t>(col_idx + Nt) + 7) & ~uintptr_t(7));
  const int smem = (int)sizeof(T) * D * (kMatchRows + kMatchTile);
  const int grid = (Ns + kMatchRows - 1) / kMatchRows;
  ensure_dynamic_smem(reinterpret_cast<const void*>(match_kernel<T, 0>), smem);
  ensure_dynamic_smem(reinterpret_cast<const void*>(match_kernel<T, 1>), smem);
  if (mutual) match_init_kernel<Bits><<<(Nt + 255) / 256, 256, 0, st>>>(col_min, col_idx, Nt);
  match_kernel<T, 0><<<grid, kMatchRows, smem, st>>>(src_desc, tgt_desc, Ns, Nt, D, row_idx, mutual ? col_min : nullptr, col_idx);
  if (mutual) match_kernel<T, 1><<<grid, kMatchRows, smem, st>>>(src_desc, tgt_desc, Ns, Nt, D, row_idx, col_min, col_idx);
  compact_center_kernel<<<1, kCompactThreads, 0, st>>>(row_idx, col_idx, src_keypts, tgt_keypts, Ns, mutual, corr, count, corr_pos,
                                                       out_src, out_tgt);
}

size_t match_scratch_bytes(int Ns, int Nt) { return (size_t)(Ns + Nt) * sizeof(int32_t) + (size_t)Nt * 8 + 16; }
int match_max_dim() { return kMatchMaxD; }

void launch_match(const void* src_desc, const void* tgt_desc, int desc_is_fp64, const float* src_keypts, const float* tgt_keypts,
                  int Ns, int Nt, int D, int mutual, void* scratch, int32_t* corr, int32_t* count, float* corr_pos,
                  float* out_src, float* out_tgt, cudaStream_t st) {
  if (desc_is_fp64)
    launch_match_t<double>(static_cast<const double*>(src_desc), static_cast<const double*>(tgt_desc), src_keypts, tgt_keypts, Ns,
                           Nt, D, mutual, scratch, corr, count, corr_pos, out_src, out_tgt, st);
  else
    launch_match_t<float>(static_cast<const float*>(src_desc), static_cast<const float*>(tgt_desc), src_keypts, tgt_keypts, Ns,
                          Nt, D, mutual, scratch, corr, count, corr_pos, out_src, out_tgt, st);
}

}  // namespace pdsc
