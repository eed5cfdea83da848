// Micro-benchmark 2: what bounds a TMA tile load on B200 — bytes in flight x latency, or the box shape?
// 148 CTAs, one elected thread each keeps `stages` boxes in flight from a tensor larger than L2 (cold HBM reads)
// or re-reads a small region (L2 hits).  Shapes:
//   A  2-D [128 rows x 128 B], rows contiguous (K = 32 floats)                — pointwise GEMM A tile, ld = 32
//   B  2-D [128 rows x 128 B], row pitch 384 B                                — pointwise GEMM A tile, ld = 96
//   C  4-D halo tile (32 ch, 34 x 6 px) of NHWC [B,52,52,96], 128B swizzle     — fused dw->pw raw tile (204 rows)
//   D  4-D halo tile (96 ch, 34 x 6 px), no swizzle, 384-byte rows            — same pixels, all channels in one box
//   E  4-D halo tile (64 ch, 34 x 6 px), no swizzle, 256-byte rows
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo_nano_b200/csrc -I include -o tools/_bin/tma_probe2 tools/tma_probe2.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace ynb;

struct Args {
  int shape;        // 0 = 2-D, 1 = 4-D halo
  int box_bytes;
  int boxes;        // boxes per CTA
  int stages;
  int hot;          // 1: every CTA re-reads the same few tiles (L2 hits)
  int tiles_x, tiles_y, TW, TH, cstep, nchunks;
};

__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap tm, const Args a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32 && ptx::elect_one()) {
    const uint32_t stage_stride = ((uint32_t)a.box_bytes + 1023u) & ~1023u;
    int issued = 0, done = 0;
    long long lat_sum = 0;
    long long t_issue[8];
    const long long t0 = clock64();
    while (done < a.boxes) {
      while (issued < a.boxes && issued - done < a.stages) {
        const int s = issued % a.stages;
        int idx = a.hot ? (issued % 4) : (blockIdx.x + issued * gridDim.x);
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)a.box_bytes);
        if (a.shape == 0) {
          ptx::tma_load_2d(smem + s * stage_stride, &tm, &full[s], 0, idx * 128);
        } else {
          const int kc = idx % a.nchunks;
          const int t = idx / a.nchunks;
          const int per_img = a.tiles_x * a.tiles_y;
          const int b = t / per_img, r = t % per_img;
          ptx::tma_load_4d(smem + s * stage_stride, &tm, &full[s], kc * a.cstep, (r % a.tiles_x) * a.TW - 1,
                           (r / a.tiles_x) * a.TH - 1, b);
        }
        t_issue[s] = clock64();
        ++issued;
      }
      const int s = done % a.stages;
      while (!ptx::mbar_try_wait(&full[s], (uint32_t)((done / a.stages) & 1))) {}
      lat_sum += clock64() - t_issue[s];
      ++done;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = lat_sum; }
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
  long long* d; cudaMalloc(&d, 16);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int B = 256, H = 52, W = 52, C = 96;                       // 266 MB > L2
  float* x; cudaMalloc(&x, (size_t)B * H * W * C * 4); cudaMemset(x, 0, (size_t)B * H * W * C * 4);
  float* y; cudaMalloc(&y, (size_t)256 << 20); cudaMemset(y, 0, (size_t)256 << 20);   // flusher / 2-D source
  struct Case { const char* name; int shape; CUtensorMap tm; int box_bytes; int cstep, nchunks; };
  std::vector<Case> cases;
  {
    Case c{"A 2-D 128x128B contiguous      ", 0, {}, 16384, 0, 1};
    cuuint64_t dims[2] = {32, (cuuint64_t)(256 << 20) / 128}; cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    enc(&c.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cases.push_back(c);
  }
  {
    Case c{"B 2-D 128x128B pitch 384B      ", 0, {}, 16384, 0, 1};
    cuuint64_t dims[2] = {96, (cuuint64_t)B * H * W}; cuuint64_t strides[1] = {384};
    cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    enc(&c.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cases.push_back(c);
  }
  auto halo = [&](const char* name, int cbox, CUtensorMapSwizzle sw) {
    Case c{name, 1, {}, cbox * 4 * 34 * 6, cbox, C / cbox};
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * 4 * W, (cuuint64_t)C * 4 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)cbox, 34, 6, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&c.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed for %s: %d\n", name, (int)r);
    else cases.push_back(c);
  };
  halo("C 4-D halo 32ch x34x6 swizzle128 ", 32, CU_TENSOR_MAP_SWIZZLE_128B);
  halo("D 4-D halo 96ch x34x6 no swizzle ", 96, CU_TENSOR_MAP_SWIZZLE_NONE);
  halo("E 4-D halo 48ch x34x6 no swizzle ", 48, CU_TENSOR_MAP_SWIZZLE_NONE);
  for (auto& c : cases)
    for (int hot : {0, 1})
      for (int stages : {1, 2, 4, 6}) {
        const int stride = (c.box_bytes + 1023) & ~1023;
        if (stages * stride > smem_bytes - 2048) continue;
        Args a{c.shape, c.box_bytes, 48, stages, hot, 2, 13, 32, 4, c.cstep, c.nchunks};
        cudaMemset(y, 1, (size_t)256 << 20);        // flush L2
        tma_kernel<<<148, 128, smem_bytes>>>(c.tm, a, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s %s stages %d: %7.0f cycles/box latency, %6.1f B/clk/SM  [%s]\n", c.name, hot ? "L2-hot " : "HBM    ", stages,
               (double)h[1] / a.boxes, (double)c.box_bytes * a.boxes / h[0], cudaGetErrorString(e));
      }
  return 0;
}
