// Developer probe (not part of the product): tcgen05.mma with A in TENSOR MEMORY and B in the MN-MAJOR shared-memory layout
// (B stored [K rows][N contiguous], i.e. V as [key][channel]); tries the candidate descriptor encodings.
//   D[128 x 128] (fp32, TMEM) = A[128 x 64] (fp16, TMEM: lane = row, 32-bit column c holds K = 2c | 2c+1) * B^T
//   with B = [128 rows (N) x 64 (K)] fp16, K-major SWIZZLE_128B panel in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I pointdsc_b200/csrc tools/ts_mma_probe.cu -o /tmp/ts_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

using namespace pdsc::ptx;

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const uint8_t* Bimg, float* D, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, uint32_t bmajor_bit) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = smem_u32(smem);
  for (int i = tid; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t tA = tmem + 128;  // A image: 32 columns (64 fp16) at +128; D at +0 (128 columns)
  // each thread writes its row of A: 32 packed columns
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) {
    const __half2 h = __halves2half2(A[tid * 64 + 2 * c], A[tid * 64 + 2 * c + 1]);
    v[c] = *reinterpret_cast<const uint32_t*>(&h);
  }
  tmem_st32(tA + ((uint32_t)(warp * 32) << 16), v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_f16kind(128, 128, 0) | (1u << bmajor_bit);
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t addr = s0 + ks * kstep_bytes;
      const uint64_t bdesc = (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
                             ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
          ::"r"(tmem), "r"(tA + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc)
          : "memory");
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t o[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, o);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[tid * 128 + c0 + i] = __uint_as_float(o[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  std::vector<__half> A(128 * 64), B(128 * 64);
  std::vector<uint8_t> img(16384, 0);
  srand(1);
  for (auto& x : A) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 64; ++k) {
      const __half h = __float2half((rand() % 2001 - 1000) / 1000.0f);
      B[n * 64 + k] = h;
      *reinterpret_cast<__half*>(img.data() + (n / 64) * 8192 + (k / 8) * 1024 + (k % 8) * 128 + ((((n % 64) / 8) ^ (k % 8)) * 16) + (n % 8) * 2) = h;
    }
  __half* dA; uint8_t* dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, img.size()); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const uint32_t cand[][4] = {  // lbo, sbo, bytes per K=16 step, idesc bit of "B is MN-major"
      {8192, 1024, 2048, 16}, {1024, 8192, 2048, 16}, {8192, 1024, 2048, 15}, {1024, 8192, 2048, 15},
      {8192, 2048, 2048, 16}, {2048, 8192, 2048, 16}, {8192, 1024, 32, 16}, {16, 1024, 2048, 16}};
  for (auto& c : cand) {
    cudaMemset(dD, 0, 128 * 128 * 4);
    probe<<<1, 128, 32768>>>(dA, dB, dD, c[0], c[1], c[2], c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo %u sbo %u kstep %u bit %u: launch failed: %s\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * 128);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)__half2float(A[m * 64 + k]) * (double)__half2float(B[n * 64 + k]);
        maxerr = fmax(maxerr, fabs(ref - D[m * 128 + n]));
      }
    printf("lbo %5u sbo %5u kstep %4u major-bit %2u: max |D - ref| = %.3e %s\n", c[0], c[1], c[2], c[3], maxerr, maxerr < 1e-3 ? "<== MATCH" : "");
  }
  return 0;
}
