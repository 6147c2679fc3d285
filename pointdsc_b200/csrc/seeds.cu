// a6 — seed selection by parallel non-maximum suppression (testing mode).
//
// Reference: models/PointDSC.py:199-217 (pick_seeds)
//   is_local_max_i = all_j ( s_i >= s_j  or  ||x_i - x_j|| >= R )
//   seeds          = argsort(s * is_local_max, descending)[:S]
// The reference reads the materialised N x N `src_dist`; here the distance is recomputed from the
// points with the same rounded operations (common.cuh::length3), so the N x N matrix never exists.
// Ranking contract (SURVEY.md §7 trap 3): keys are compared as fp32 values with +0 == -0, and exact
// ties are broken by the LOWEST index (a stable descending sort; the reference's argsort is unstable,
// so any order of tied keys is a valid reference output).
#include "common.cuh"
#include "kernels.h"

namespace pdsc {

constexpr int kSeedMaxN = 16384;

__global__ void __launch_bounds__(256) nms_key_kernel(const float* __restrict__ src, const float* __restrict__ conf,
                                                      float* __restrict__ key, int N, float radius) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  if (i >= N) return;
  const float* p = src + (size_t)b * N * 3;
  const float* s = conf + (size_t)b * N;
  const float si = s[i];
  const float xi = p[(size_t)i * 3], yi = p[(size_t)i * 3 + 1], zi = p[(size_t)i * 3 + 2];
  bool ok = true;
  for (int j = lane; j < N; j += 32) {
    const float d = length3(xi - p[(size_t)j * 3], yi - p[(size_t)j * 3 + 1], zi - p[(size_t)j * 3 + 2]);
    ok = ok && ((si >= s[j]) || (d >= radius));
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) key[(size_t)b * N + i] = si * (ok ? 1.0f : 0.0f);
}

__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) == 0u) u = 0u;  // -0 ranks equal to +0, as in a floating-point comparison
  return u ^ ((u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}

// one CTA per set: bitonic sort of (descending key, ascending index), emit the first S indices
__global__ void __launch_bounds__(1024) seed_sort_kernel(const float* __restrict__ key, int32_t* __restrict__ seeds,
                                                         int N, int P, int S) {
  extern __shared__ unsigned long long skeys[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long v = ~0ull;
    if (i < N) v = ((unsigned long long)(~orderable(key[(size_t)b * N + i])) << 32) | (unsigned)i;
    skeys[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skeys[i], c = skeys[ixj];
          const bool asc = (i & k) == 0;
          if ((a > c) == asc) { skeys[i] = c; skeys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) seeds[(size_t)b * S + i] = (int32_t)(skeys[i] & 0xFFFFFFFFull);
}

int pick_seeds_max_n() { return kSeedMaxN; }

void launch_pick_seeds(const float* src, const float* conf, int32_t* seeds, float* key_scratch, int B, int N, int S,
                       float radius, cudaStream_t st) {
  dim3 g1((N + 7) / 8, B);
  nms_key_kernel<<<g1, 256, 0, st>>>(src, conf, key_scratch, N, radius);
  int P = 2;
  while (P < N) P <<= 1;
  const int smem = P * (int)sizeof(unsigned long long);
  static int configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(seed_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeedMaxN * 8);
    configured = kSeedMaxN * 8;
  }
  const int threads = P / 2 < 1024 ? (P / 2 < 32 ? 32 : P / 2) : 1024;
  seed_sort_kernel<<<B, threads, smem, st>>>(key_scratch, seeds, N, P, S);
}

}  // namespace pdsc
