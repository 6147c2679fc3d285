// Warp-level "smallest `want` of N keys, sorted" on order-preserving 32-bit keys in shared memory (nsm.cu: feature-space kNN of
// the seed rows; fpfh.cu: hybrid radius / k-nearest search).  See the call sites for the key encodings.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace pdsc {

// float -> order-preserving uint32 (+0 == -0, NaN last)
__device__ __forceinline__ uint32_t dist_key32(float d) {
  uint32_t u = __float_as_uint(d);
  if (d != d) return 0xFFFFFFFFu;            // NaN distances are never selected
  if ((u & 0x7FFFFFFFu) == 0u) u = 0u;       // -0 ranks equal to +0
  return u ^ ((u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}

// keys[0..NP) (NP = N rounded up to 32; entries >= N are ignored), hist[256], sel[P] (P = power of two >= want) are this warp's
// shared memory.  On return sel[0..want) holds the `want` smallest (key << 32 | index) pairs in ascending order — ties by
// ascending index — and sel[want..P) is ~0.  want <= N.  A 4-pass radix SELECT (8 bits per pass, 256-bin histogram) finds the
// value T of the want-th smallest key and how many elements equal to T belong to the selection; one ordered pass compacts the
// winners (every key < T plus the lowest-index elements with key == T); a bitonic sort orders them.
__device__ __forceinline__ void warp_select_sorted(uint32_t* keys, uint32_t* hist, unsigned long long* sel, int N, int NP, int want,
                                                   int P, int lane) {
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  for (int j = lane; j < N; j += 32) {
    const uint32_t key = keys[j];
    kmin = min(kmin, key); kmax = max(kmax, key);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  const uint32_t lt_mask = (1u << lane) - 1u;
  // ---- radix select of the (k+1)-th smallest key -----------------------------------------------------
  // Distances of unit vectors share sign and most exponent bits: the digits START at the highest bit in which the row's keys
  // differ (everything above it is common), so the first histogram already spreads over its 256 bins.  Starting at bit 31 the
  // first two passes put nearly every key of a row into one or two bins — 32-way serialised shared-memory atomics, which made
  // this kernel 1.6 ms of the KITTI N = 5000 configuration.
  uint32_t prefix = 0u, T;
  int need = want;                          // rank (1-based) still to be located among the keys matching `prefix`
  if (kmin == kmax) {
    T = kmin;                                // every key equal: the lowest indices
  } else {
    int hi_bit = 31 - __clz(kmin ^ kmax);    // highest differing bit
    prefix = (hi_bit == 31) ? 0u : (kmin >> (hi_bit + 1)) << (hi_bit + 1);
#pragma unroll 1
    while (hi_bit >= 0) {
      const int width = hi_bit + 1 < 8 ? hi_bit + 1 : 8;
      const int shift = hi_bit + 1 - width;
      const uint32_t dmask = (1u << width) - 1u;
#pragma unroll
      for (int q = 0; q < 8; ++q) hist[lane * 8 + q] = 0u;
      __syncwarp();
      for (int j = lane; j < NP; j += 32) {
        const uint32_t v = keys[j];
        const bool in = (hi_bit == 31) || ((v >> (hi_bit + 1)) == (prefix >> (hi_bit + 1)));
        if (in && j < N) atomicAdd(&hist[(v >> shift) & dmask], 1u);
      }
      __syncwarp();
      uint32_t c[8], lane_sum = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = hist[lane * 8 + q]; lane_sum += c[q]; }
      uint32_t incl = lane_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t excl = incl - lane_sum;
      const bool mine = excl < (uint32_t)need && (uint32_t)need <= incl;   // exactly one lane (need <= matching count)
      uint32_t digit = 0u, rem = 0u;
      if (mine) {
        uint32_t cum = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (cum < (uint32_t)need && (uint32_t)need <= cum + c[q]) { digit = (uint32_t)(lane * 8 + q); rem = (uint32_t)need - cum; }
          cum += c[q];
        }
      }
      const int srcl = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
      digit = __shfl_sync(0xffffffffu, digit, srcl);
      need = (int)__shfl_sync(0xffffffffu, rem, srcl);
      prefix |= digit << shift;
      hi_bit = shift - 1;
      __syncwarp();
    }
    T = prefix;
  }
  // ---- ordered compaction: key < T, plus the `need` lowest indices with key == T ----------------------
  int out = 0, eq_seen = 0;
  for (int j0 = 0; j0 < NP; j0 += 32) {
    const uint32_t v = keys[j0 + lane];
    const bool eq = v == T && (j0 + lane < N);
    const uint32_t beq = __ballot_sync(0xffffffffu, eq);
    const bool take = (j0 + lane < N) && ((v < T) || (eq && eq_seen + __popc(beq & lt_mask) < need));
    const uint32_t bt = __ballot_sync(0xffffffffu, take);
    if (take) sel[out + __popc(bt & lt_mask)] = ((unsigned long long)v << 32) | (unsigned)(j0 + lane);
    out += __popc(bt);
    eq_seen += __popc(beq);
  }
  for (int i = out + lane; i < P; i += 32) sel[i] = ~0ull;      // out == want <= P
  __syncwarp();
  // ---- bitonic sort of the P packed pairs (ascending) -------------------------------------------------
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));   // the lower index of the t-th pair at distance jj
        const int ixj = i | jj;
        const unsigned long long a = sel[i], c2 = sel[ixj];
        const bool asc = (i & kk) == 0;
        if ((a > c2) == asc) { sel[i] = c2; sel[ixj] = a; }
      }
      __syncwarp();
    }
  }
}

}  // namespace pdsc
