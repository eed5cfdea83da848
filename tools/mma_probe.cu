// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) on sm_100a,
// operands A and B in shared memory (SS), one CTA per SM, MMAs issued back to back by one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I yolo_nano_b200/csrc -I include -o mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace ynb;

__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int iters, int nacc, long long* out, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar2[i], 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc(2, 128, N);
    const uint32_t a = ptx::smem_u32(smem), b = ptx::smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t ko = (i & 3) * 32;
      ptx::mma_tf32_ss(tmem + (uint32_t)((i % nacc) * N), ptx::make_sw128_kmajor_desc(a + ko),
                       ptx::make_sw128_kmajor_desc(b + ko), idesc, i >= nacc);
      if (commit_every > 0 && (i + 1) % commit_every == 0) ptx::mma_commit(&bar2[(i / commit_every) & 3]);   // nobody waits
    }
    ptx::mma_commit(&bar);
    while (!ptx::mbar_try_wait(&bar, 0)) {}
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  for (int grid : {1, 148}) {
    for (int N : {64, 96, 128, 256}) {
      for (int nacc : {1, 2}) {
        if (nacc * N > 512) continue;
        mma_rate_kernel<<<grid, 128, 64 * 1024>>>(N, iters, nacc, d, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cyc / iters;
        printf("grid %3d  N=%3d  accumulators %d: %7.1f cycles per MMA (128x%dx8)  = %6.1f TFLOP/s over 148 SMs @1.9 GHz  [%s]\n",
               grid, N, nacc, per, N, 2.0 * 128 * N * 8 / per * 148 * 1.9e9 / 1e12, cudaGetErrorString(e));
      }
    }
  }
  for (int ce : {0, 16, 8, 4, 2, 1}) {
    mma_rate_kernel<<<148, 128, 64 * 1024>>>(128, iters, 2, d, ce);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("N=128, tcgen05.commit every %2d MMAs: %7.1f cycles per MMA  [%s]\n", ce, (double)cyc / iters, cudaGetErrorString(e));
  }
  return 0;
}
