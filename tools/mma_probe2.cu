// Micro-benchmark 2: issue cost of tcgen05.mma by kind (tf32 / bf16), M (64 / 128) and N, operands in
// shared memory, one CTA per SM, MMAs issued back to back by one elected thread of a warp-uniform branch.
// Answers: is the per-instruction cost flat in N for kind::f16 too (i.e. what does the bf16 mode buy)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo_nano_b200/csrc -I include \
//        -o tools/_bin/mma_probe2 tools/mma_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace ynb;

template <int KIND>   // 0: tf32, 1: bf16 (kind::f16)
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) mma_rate_kernel(int M, int N, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x < 32) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc(KIND == 0 ? 2 : 1, M, N);
      const uint32_t a = ptx::smem_u32(smem), b = ptx::smem_u32(smem + 16384);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t ko = (i & 3) * 32;
        mma_ss<KIND>(tmem + (uint32_t)((i % nacc) * N), ptx::make_sw128_kmajor_desc(a + ko),
                     ptx::make_sw128_kmajor_desc(b + ko), idesc, i >= nacc);
      }
      ptx::mma_commit(&bar);
      while (!ptx::mbar_try_wait(&bar, 0)) {}
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  for (int kind : {0, 1})
    for (int M : {128, 64})
      for (int N : {16, 32, 64, 96, 128, 192, 256})
        for (int nacc : {1, 2}) {
          if (nacc * N > 512) continue;
          if (M == 128 && N % 16) continue;
          if (kind == 0) mma_rate_kernel<0><<<148, 128, 64 * 1024>>>(M, N, iters, nacc, d);
          else mma_rate_kernel<1><<<148, 128, 64 * 1024>>>(M, N, iters, nacc, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
          const double per = (double)cyc / iters;
          const int K = kind == 0 ? 8 : 16;
          printf("%s M=%3d N=%3d K=%2d acc %d: %7.1f cycles per MMA = %7.1f TFLOP/s (148 SMs @1.9 GHz) [%s]\n",
                 kind == 0 ? "tf32" : "bf16", M, N, K, nacc, per, 2.0 * M * N * K / per * 148 * 1.9e9 / 1e12,
                 cudaGetErrorString(e));
        }
  return 0;
}
