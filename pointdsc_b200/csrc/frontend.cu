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
//   match_rows_kernel<T>   nearest column of every row of a [rows x cols] problem over ONE chunk of the columns (grid.y chunks, so
//                          that a 5000 x 5000 problem fills the GPU): thread = row, its descriptor cached 16 channels at a time in
//                          registers; the chunk's descriptors are staged channel-major in shared memory, 64 columns (fp32) or 32
//                          (fp64) per tile, and read as BROADCAST 128-bit loads of four (two) adjacent columns — 64 FMAs per 16
//                          loads, every dot product one ascending-channel FMA chain in T.  The square root is applied BEFORE the
//                          strict '<' comparison, as the reference does (it merges distances that differ by less than an ulp of
//                          the root), which is numpy's first-minimum rule inside the chunk.  Result: (distance, index) per
//                          (row, chunk).
//   match_reduce_kernel<T> thread = row: scans its chunks in ascending column order with the same strict '<' — the first
//                          minimum of the whole row.  No atomics anywhere.
//   The mutual check needs argmin along the other axis: the SAME two kernels with the roles of the two descriptor sets
//   swapped.  a*b == b*a exactly and the channel order is the same, so both launches see bit-identical distances — the
//   consistency numpy gets from taking both argmins of one matrix.
//   compact_center_kernel  one CTA: ordered compaction of the kept pairs (ascending source index), gather of the key points,
//                          column means in fp64, centred corr_pos — written in the layout pdsc_forward consumes.
// The reference's BLAS fixes no accumulation order, so indices are only comparable where the best and the second-best distance
// are separated (tests); exact duplicates resolve to the first one, as in numpy.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace pdsc {

constexpr int kMatchRows = 128;   // rows per CTA (one per thread)
constexpr int kMatchMaxD = 64;
constexpr int kMatchMaxChunks = 32;

template <typename T>
__device__ __forceinline__ T feature_distance(T dot) {
  // np.sqrt(2 - 2 * dot + 1e-6): python scalars adopt the array dtype, every operation rounded separately
  if (sizeof(T) == 4) return (T)__fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, (float)dot)), 1e-6f));
  return (T)__dsqrt_rn(__dadd_rn(__dsub_rn(2.0, __dmul_rn(2.0, (double)dot)), 1e-6));
}
template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c) {
  if (sizeof(T) == 4) return (T)__fmaf_rn((float)a, (float)b, (float)c);
  return (T)__fma_rn((double)a, (double)b, (double)c);
}

template <typename T> struct MatchPartial { T dist; int idx; int pad; };

template <typename T>
__global__ void __launch_bounds__(kMatchRows) match_rows_kernel(const T* __restrict__ row_desc, const T* __restrict__ col_desc,
                                                                int rows, int cols, int D, int cols_per_chunk,
                                                                MatchPartial<T>* __restrict__ partial) {
  constexpr int TT = sizeof(T) == 4 ? 64 : 32;        // columns per shared-memory tile
  constexpr int VW = 16 / (int)sizeof(T);             // columns per 128-bit load
  extern __shared__ __align__(16) unsigned char match_smem[];
  T* s_row = reinterpret_cast<T*>(match_smem);        // [D][kMatchRows]  channel-major: conflict-free per-thread reads
  T* s_col = s_row + (size_t)D * kMatchRows;          // [D][TT]          channel-major: broadcast vector reads
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kMatchRows + tid;
  const int c_begin = blockIdx.y * cols_per_chunk, c_end = min(cols, c_begin + cols_per_chunk);
  for (int e = tid; e < D * kMatchRows; e += kMatchRows) {
    const int r = e / D, d = e % D;                   // coalesced over the row-major descriptor block
    const int gi = blockIdx.x * kMatchRows + r;
    s_row[(size_t)d * kMatchRows + r] = gi < rows ? row_desc[(size_t)gi * D + d] : (T)0;
  }
  T best = (T)0;
  int best_j = -1;
  for (int j0 = c_begin; j0 < c_end; j0 += TT) {
    __syncthreads();
    const int cnt = min(TT, c_end - j0);
    for (int e = tid; e < TT * D; e += kMatchRows) {
      const int t = e / D, d = e % D;
      s_col[(size_t)d * TT + t] = t < cnt ? col_desc[(size_t)(j0 + t) * D + d] : (T)0;
    }
    __syncthreads();
    T acc[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) acc[t] = (T)0;
    for (int d0 = 0; d0 < D; d0 += 16) {
      T sv[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) sv[q] = (d0 + q < D) ? s_row[(size_t)(d0 + q) * kMatchRows + tid] : (T)0;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (d0 + q < D) {                            // warp-uniform
          const T* cv = s_col + (size_t)(d0 + q) * TT;
#pragma unroll
          for (int t = 0; t < TT; t += VW) {
            if (sizeof(T) == 4) {
              const float4 v = *reinterpret_cast<const float4*>(cv + t);
              acc[t] = fma_t<T>(sv[q], (T)v.x, acc[t]); acc[t + 1] = fma_t<T>(sv[q], (T)v.y, acc[t + 1]);
              acc[t + 2] = fma_t<T>(sv[q], (T)v.z, acc[t + 2]); acc[t + 3] = fma_t<T>(sv[q], (T)v.w, acc[t + 3]);
            } else {
              const double2 v = *reinterpret_cast<const double2*>(cv + t);
              acc[t] = fma_t<T>(sv[q], (T)v.x, acc[t]); acc[t + 1] = fma_t<T>(sv[q], (T)v.y, acc[t + 1]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      if (t < cnt) {
        const T dist = feature_distance<T>(acc[t]);
        if (best_j < 0 || dist < best) { best = dist; best_j = j0 + t; }      // strict '<': the first minimum wins
      }
    }
  }
  if (i < rows) {
    MatchPartial<T>& o = partial[(size_t)i * gridDim.y + blockIdx.y];
    o.dist = best;
    o.idx = best_j;
  }
}

template <typename T>
__global__ void match_reduce_kernel(const MatchPartial<T>* __restrict__ partial, int rows, int chunks, int32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  T best = (T)0;
  int best_j = -1;
  for (int c = 0; c < chunks; ++c) {                  // ascending column ranges: strict '<' keeps the first minimum
    const MatchPartial<T> p = partial[(size_t)i * chunks + c];
    if (p.idx >= 0 && (best_j < 0 || p.dist < best)) { best = p.dist; best_j = p.idx; }
  }
  idx[i] = best_j;
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

static int match_chunks(int rows) {
  const int row_ctas = (rows + kMatchRows - 1) / kMatchRows;
  int chunks = (4 * device_sm_count() + row_ctas - 1) / row_ctas;      // about four CTAs per SM
  return chunks < 1 ? 1 : (chunks > kMatchMaxChunks ? kMatchMaxChunks : chunks);
}

template <typename T>
static void nearest_columns(const T* row_desc, const T* col_desc, int rows, int cols, int D, MatchPartial<T>* partial, int32_t* idx,
                            cudaStream_t st) {
  constexpr int TT = sizeof(T) == 4 ? 64 : 32;
  int chunks = match_chunks(rows);
  int per = (cols + chunks - 1) / chunks;
  per = (per + TT - 1) / TT * TT;                      // whole tiles per chunk
  chunks = (cols + per - 1) / per;
  const int smem = (int)sizeof(T) * D * (kMatchRows + TT);
  ensure_dynamic_smem(reinterpret_cast<const void*>(match_rows_kernel<T>), smem);
  match_rows_kernel<T><<<dim3((rows + kMatchRows - 1) / kMatchRows, chunks), kMatchRows, smem, st>>>(row_desc, col_desc, rows, cols,
                                                                                                    D, per, partial);
  match_reduce_kernel<T><<<(rows + 255) / 256, 256, 0, st>>>(partial, rows, chunks, idx);
}

template <typename T>
static void launch_match_t(const T* src_desc, const T* tgt_desc, const float* src_keypts, const float* tgt_keypts, int Ns, int Nt,
                           int D, int mutual, void* scratch, int32_t* corr, int32_t* count, float* corr_pos, float* out_src,
                           float* out_tgt, cudaStream_t st) {
  // scratch: row_idx [Ns] int32 | col_idx [Nt] int32 | partial [max(Ns, Nt)][kMatchMaxChunks] (16-byte aligned)
  int32_t* row_idx = static_cast<int32_t*>(scratch);
  int32_t* col_idx = row_idx + Ns;
  auto* partial = reinterpret_cast<MatchPartial<T>*>((reinterpret_cast<uintptr_t>(col_idx + Nt) + 15) & ~uintptr_t(15));
  nearest_columns<T>(src_desc, tgt_desc, Ns, Nt, D, partial, row_idx, st);
  if (mutual) nearest_columns<T>(tgt_desc, src_desc, Nt, Ns, D, partial, col_idx, st);   // argmin along the other axis
  compact_center_kernel<<<1, kCompactThreads, 0, st>>>(row_idx, col_idx, src_keypts, tgt_keypts, Ns, mutual, corr, count, corr_pos,
                                                       out_src, out_tgt);
}

size_t match_scratch_bytes(int Ns, int Nt) {
  const size_t rows = (size_t)(Ns > Nt ? Ns : Nt);
  return (size_t)(Ns + Nt) * sizeof(int32_t) + rows * kMatchMaxChunks * sizeof(MatchPartial<double>) + 32;
}
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
