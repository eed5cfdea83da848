// Micro-benchmark: how fast can ONE SM pull [rows x 128 B] boxes through TMA (cp.async.bulk.tensor.2d)?
// A ring of smem stages, one thread issues, the same thread waits; no MMA, no consumers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo_nano_b200/csrc -I include -o tools/_bin/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace ynb;

__global__ void __launch_bounds__(128) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int k_chunks,
                                                       int tiles_per_cta, int stages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_all[4][8];
  if (threadIdx.x == 0) {
    for (int t = 0; t < 4; ++t)
      for (int s = 0; s < 8; ++s) ptx::mbar_init(&full_all[t][s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // `issuers` threads (one per warp) each run an independent ring: does the per-box cost overlap across threads?
  const int issuers = tiles_per_cta >> 16;
  tiles_per_cta &= 0xffff;
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < issuers) {
    uint64_t* full = full_all[threadIdx.x >> 5];
    smem += (threadIdx.x >> 5) * 32768;
    const int total = tiles_per_cta * k_chunks;
    const uint32_t bytes = (uint32_t)box_rows * 128;
    int issued = 0, done = 0;
    const long long t0 = clock64();
    while (done < total) {
      while (issued < total && issued - done < stages) {
        const int s = issued % stages;
        const int tile = blockIdx.x + (issued / k_chunks) * gridDim.x, kc = issued % k_chunks;
        ptx::mbar_arrive_expect_tx(&full[s], bytes);
        ptx::tma_load_2d(smem + (s & 1) * 16384, &tm, &full[s], kc * 32, (tile * 4 + (int)(threadIdx.x >> 5)) * box_rows);
        ++issued;
      }
      const int s = done % stages;
      while (!ptx::mbar_try_wait(&full[s], (uint32_t)((done / stages) & 1))) {}
      ++done;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 2048);
  const int rows_total = 148 * 128 * 16;
  for (int K : {128}) {
    float* x; cudaMalloc(&x, (size_t)rows_total * K * 4); cudaMemset(x, 0, (size_t)rows_total * K * 4);
    for (int box_rows : {128, 32}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows_total};
      cuuint64_t strides[1] = {(cuuint64_t)K * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int grid : {1, 148}) {
        for (int issuers : {1, 2, 4}) {
          const int k_chunks = K / 32, tiles = 4, stages = 2;
          for (int rep = 0; rep < 2; ++rep)
            tma_rate_kernel<<<grid, 128, 8 * 16384 + 2048>>>(tm, box_rows, k_chunks, tiles | (issuers << 16), stages, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
          const double boxes = (double)tiles * k_chunks;
          printf("box %3d rows grid %3d issuers %d: %7.0f cycles per box per issuer, %5.1f B/clk/SM total  [%s]\n", box_rows, grid,
                 issuers, cyc / boxes, issuers * box_rows * 128.0 * boxes / cyc, cudaGetErrorString(e));
        }
      }
    }
    cudaFree(x);
  }
  return 0;
}
