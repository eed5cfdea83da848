// Micro-benchmark: fp32 FMA issue rates on sm_100a (register, constant/uniform operand, packed FFMA2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
struct W { float w[64]; };
constexpr int ITERS = 2048;
__global__ void k_reg(float* out, float a, float b) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_const(float* out, const __grid_constant__ W wt, float a) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = a + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(x, wt.w[i + (it & 3) * 16], acc[i]);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
  float2 x = make_float2(a + threadIdx.x, a), y = make_float2(b, b + 1);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lds(float* out, float a) {
  __shared__ float4 sw[64];
  if (threadIdx.x < 64) sw[threadIdx.x] = make_float4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = a + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 w = sw[(it & 15) * 4 + i];
      acc[4 * i] = fmaf(x, w.x, acc[4 * i]); acc[4 * i + 1] = fmaf(x, w.y, acc[4 * i + 1]);
      acc[4 * i + 2] = fmaf(x, w.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(x, w.w, acc[4 * i + 3]);
    }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  W w; for (int i = 0; i < 64; ++i) w.w[i] = 1.0f + i * 1e-3f;
  const double fma = 148.0 * 8 * 256 * ITERS * 16;
  for (int threads : {256, 1024}) {
    int blocks = 148 * 8 * 256 / threads;
    double sc = 1.0;
    float t;
    t = timeit([&] { k_reg<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("threads %4d  FFMA reg      %.3f ms  %.1f TFMA/s  %.1f FMA/clk/SM@1.9GHz\n", threads, t, fma / t / 1e9, fma / (t * 1e-3) / 148 / 1.9e9);
    t = timeit([&] { k_const<<<blocks, threads>>>(out, w, 1.0001f); });
    printf("threads %4d  FFMA const/UR %.3f ms  %.1f TFMA/s  %.1f FMA/clk/SM\n", threads, t, fma / t / 1e9, fma / (t * 1e-3) / 148 / 1.9e9);
    t = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("threads %4d  FFMA2 reg     %.3f ms  %.1f TFMA/s  %.1f FMA/clk/SM\n", threads, t, 2 * fma / t / 1e9, 2 * fma / (t * 1e-3) / 148 / 1.9e9);
    t = timeit([&] { k_lds<<<blocks, threads>>>(out, 1.0001f); });
    printf("threads %4d  FFMA+LDS.128  %.3f ms  %.1f TFMA/s  %.1f FMA/clk/SM\n", threads, t, fma / t / 1e9, fma / (t * 1e-3) / 148 / 1.9e9);
    (void)sc;
  }
  return 0;
}
