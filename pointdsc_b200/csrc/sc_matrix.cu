// Stage i — the N x N second-order spatial-consistency matrix.
//
// Reference: models/PointDSC.py:150-153
//   src_dist = ||x_i - x_j||,  SC_ij = max(0, 1 - (src_dist_ij - ||y_i - y_j||)^2 / sigma_d^2)
// The reference materialises two [N,N,3] broadcast temporaries (>= 40 N^2 bytes of traffic); this
// kernel reads 24 B per point and writes each SC element exactly once (4 N NS bytes), which is the
// stage's algorithmic HBM traffic: it is write-bandwidth bound.
//
// Layout: sc[b][i][j], row stride NS = N rounded up to 64 floats so that every 64-key tile of a row
// is a 256-byte aligned segment for the attention kernels; the pad columns j >= N are written as 0.
// `src_dist` is NOT materialised: its only consumer, the seed NMS (a6), recomputes it from the points.
#include "common.cuh"
#include "kernels.h"

namespace pdsc {

constexpr int kScRows = 32;

__global__ void __launch_bounds__(256) sc_matrix_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                                                        float* __restrict__ sc, int N, int NS, float s2) {
  __shared__ float rs[kScRows][3], rt[kScRows][3];
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kScRows;
  const float* ps = src + (size_t)b * N * 3;
  const float* pt = tgt + (size_t)b * N * 3;
  for (int t = threadIdx.x; t < kScRows * 3; t += blockDim.x) {
    const int r = t / 3, c = t % 3;
    const int i = min(i0 + r, N - 1);
    rs[r][c] = ps[(size_t)i * 3 + c];
    rt[r][c] = pt[(size_t)i * 3 + c];
  }
  __syncthreads();
  const int rows = min(kScRows, N - i0);
  float* out = sc + ((size_t)b * N + i0) * NS;
  for (int j = threadIdx.x; j < NS; j += blockDim.x) {
    if (j < N) {
      const float sx = ps[(size_t)j * 3], sy = ps[(size_t)j * 3 + 1], sz = ps[(size_t)j * 3 + 2];
      const float tx = pt[(size_t)j * 3], ty = pt[(size_t)j * 3 + 1], tz = pt[(size_t)j * 3 + 2];
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float ds = length3(rs[r][0] - sx, rs[r][1] - sy, rs[r][2] - sz);
        const float dt = length3(rt[r][0] - tx, rt[r][1] - ty, rt[r][2] - tz);
        out[(size_t)r * NS + j] = consistency(__fsub_rn(ds, dt), s2);
      }
    } else {
      for (int r = 0; r < rows; ++r) out[(size_t)r * NS + j] = 0.0f;
    }
  }
}

void launch_sc_matrix(const float* src, const float* tgt, float* sc, int B, int N, int NS, float sigma_d,
                      cudaStream_t st) {
  const float s2 = sigma_d * sigma_d;  // fp32 product, as `self.sigma_spat ** 2`
  dim3 grid((N + kScRows - 1) / kScRows, B);
  sc_matrix_kernel<<<grid, 256, 0, st>>>(src, tgt, sc, N, NS, s2);
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_u32(uint32_t* p, uint32_t v, long long n, cudaStream_t st) {
  if (n > 0) fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
}
void launch_fill_u64(unsigned long long* p, unsigned long long v, long long n, cudaStream_t st) {
  if (n > 0) fill_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
}

}  // namespace pdsc
