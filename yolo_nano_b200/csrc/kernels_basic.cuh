// HBM-bound kernels of the path: stem conv + max-pool, depthwise 3x3, pass-through
// interleave (channel shuffle as a store permutation) and layout conversion for taps.
#pragma once
#include "common.cuh"

namespace ynb {

// =====================================================================================
// Stem: Conv2d(3,24,3,s2,p1, no bias)+BN(folded)+ReLU -> MaxPool2d(3,s2,p1), one kernel.
// (backbone/shufflenetv2.py:109-116,158-159)
//   x   NCHW [B,3,S,S]      out NHWC [B,S/4,S/4,24]
//   w   [27][24], index (ci*9 + ky*3 + kx)*24 + co     b [24]
// A CTA owns an 8x8 tile of pooled outputs: 17x17 conv outputs, 35x35x3 input pixels.
// The 24x208x208 conv map (the largest tensor of the network) never reaches HBM.
// MaxPool pads with -inf; conv outputs are post-ReLU (>= 0) and every pooling window holds
// at least one valid element, so padding with 0 gives the same maximum.
// =====================================================================================
constexpr int kStemTile = 8;
constexpr int kStemConv = 2 * kStemTile + 1;   // 17
constexpr int kStemIn = 2 * kStemConv + 1;     // 35
constexpr int kStemInPitch = 36;
constexpr int kStemC = 24;
constexpr int kStemThreads = 320;              // 289 conv positions of a tile, one per thread

// Folded stem weights travel as a kernel parameter: they then live in the constant bank
// and every FFMA takes its weight operand straight from c[0][..] (warp-uniform, no
// shared-memory traffic) — the conv is FMA-bound instead of LDS-bound.
struct StemWeights {
  float w[27][kStemC];   // [(ci*3+ky)*3+kx][co]
  float b[kStemC];
};

__global__ void __launch_bounds__(kStemThreads)
stem_pool_kernel(const float* __restrict__ x, float* __restrict__ out, const __grid_constant__ StemWeights wt,
                 int S) {
  __shared__ float s_in[3][kStemIn][kStemInPitch];
  __shared__ float s_conv[kStemConv * kStemConv][kStemC + 1];

  const int Hc = S / 2, Hp = S / 4;
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * kStemTile, px0 = blockIdx.x * kStemTile;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;   // first conv row / col of the tile
  const int iy0 = 2 * cy0 - 1, ix0 = 2 * cx0 - 1;   // first input row / col
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();          // the previous forward may still be reading / writing these buffers

  const float* xb = x + (size_t)b * 3 * S * S;
  for (int i = tid; i < 3 * kStemIn * kStemIn; i += kStemThreads) {
    int c = i / (kStemIn * kStemIn);
    int r = (i / kStemIn) % kStemIn;
    int q = i % kStemIn;
    int iy = iy0 + r, ix = ix0 + q;
    float v = 0.0f;
    if (iy >= 0 && iy < S && ix >= 0 && ix < S) v = __ldg(xb + ((size_t)c * S + iy) * S + ix);
    s_in[c][r][q] = v;
  }
  __syncthreads();

  // conv + bias + ReLU: one thread per conv position, all 24 output channels in registers
  if (tid < kStemConv * kStemConv) {
    const int r = tid / kStemConv, q = tid % kStemConv;
    const int cy = cy0 + r, cx = cx0 + q;
    float acc[kStemC];
#pragma unroll
    for (int co = 0; co < kStemC; ++co) acc[co] = wt.b[co];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = s_in[ci][2 * r + ky][2 * q + kx];
#pragma unroll
          for (int co = 0; co < kStemC; ++co) acc[co] = fmaf(v, wt.w[(ci * 3 + ky) * 3 + kx][co], acc[co]);
        }
    const bool inside = cy >= 0 && cy < Hc && cx >= 0 && cx < Hc;
#pragma unroll
    for (int co = 0; co < kStemC; ++co) s_conv[tid][co] = inside ? fmaxf(acc[co], 0.0f) : 0.0f;
  }
  __syncthreads();

  for (int i = tid; i < kStemTile * kStemTile * kStemC; i += kStemThreads) {
    int co = i % kStemC;
    int p = i / kStemC;
    int r = p / kStemTile, q = p % kStemTile;
    int py = py0 + r, px = px0 + q;
    if (py >= Hp || px >= Hp) continue;
    float m = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        m = fmaxf(m, s_conv[(2 * r + dy) * kStemConv + 2 * q + dx][co]);
    out[(((size_t)b * Hp + py) * Hp + px) * kStemC + co] = m;
  }
}

// w: [27][24] device or host floats are copied into the parameter block by the caller.
inline cudaError_t launch_stem_pool(const float* x, float* out, const StemWeights& wt, int batch, int S,
                                    cudaStream_t st) {
  int Hp = S / 4;
  dim3 grid((Hp + kStemTile - 1) / kStemTile, (Hp + kStemTile - 1) / kStemTile, batch);
  cudaError_t r = launch_pdl(stem_pool_kernel, grid, dim3(kStemThreads), 0, st, x, out, wt, S);
  YNB_COUNT_LAUNCH();
  return r;
}

// =====================================================================================
// Depthwise 3x3, pad 1, stride 1|2, + bias (+ activation).  NHWC, 4 channels per thread
// with 16-byte loads; consecutive threads take consecutive channel groups of one pixel,
// so a warp reads whole 128-byte lines.  (backbone/shufflenetv2.py:66-67; heads
// models/yolo_nano.py:51,53)
//   w [9][C4] tap-major, b [C4]  (C4 = channels rounded to 4; pads are zero)
// =====================================================================================
// Each thread produces kDwTX consecutive output pixels of one 4-channel group: the 3 x
// ((TX-1)*stride+3) input window is loaded once into registers (4.5 instead of 9 16-byte
// loads per output at stride 1) and the 9 filter taps once per thread, which takes the
// kernel off the L1 bandwidth limit the one-pixel-per-thread version sat on.
constexpr int kDwTX = 4;

template <int STRIDE>
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const float* __restrict__ in, int in_ld, int in_off,
                 float* __restrict__ out, int out_ld, int out_off,
                 const float* __restrict__ w, const float* __restrict__ bias,
                 int batch, int Hin, int Win, int C4, int act) {
  constexpr int NIN = (kDwTX - 1) * STRIDE + 3;
  const int Ho = (Hin - 1) / STRIDE + 1, Wo = (Win - 1) / STRIDE + 1;
  const int groups = C4 >> 2;
  const int xgroups = (Wo + kDwTX - 1) / kDwTX;
  const int64_t total = (int64_t)batch * Ho * xgroups * groups;
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    int64_t p = i / groups;
    const int xg = (int)(p % xgroups);
    const int yo = (int)((p / xgroups) % Ho);
    const int b = (int)(p / ((int64_t)xgroups * Ho));
    const int c = g << 2;
    const int xo0 = xg * kDwTX;
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 acc[kDwTX];
#pragma unroll
    for (int t = 0; t < kDwTX; ++t) acc[t] = bv;
    const float* inb = in + (size_t)b * Hin * Win * in_ld + in_off + c;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yi = yo * STRIDE + ky - 1;
      if (yi < 0 || yi >= Hin) continue;
      float4 v[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const int xi = xo0 * STRIDE - 1 + j;
        v[j] = (xi >= 0 && xi < Win) ? __ldg(reinterpret_cast<const float4*>(inb + ((size_t)yi * Win + xi) * in_ld))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 k = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C4 + c));
#pragma unroll
        for (int t = 0; t < kDwTX; ++t) {
          const float4 u = v[t * STRIDE + kx];
          acc[t].x = fmaf(u.x, k.x, acc[t].x);
          acc[t].y = fmaf(u.y, k.y, acc[t].y);
          acc[t].z = fmaf(u.z, k.z, acc[t].z);
          acc[t].w = fmaf(u.w, k.w, acc[t].w);
        }
      }
    }
    float* ob = out + (((size_t)b * Ho + yo) * Wo + xo0) * out_ld + out_off + c;
#pragma unroll
    for (int t = 0; t < kDwTX; ++t) {
      if (xo0 + t < Wo) {
        float4 a = acc[t];
        a.x = apply_act(a.x, act); a.y = apply_act(a.y, act); a.z = apply_act(a.z, act); a.w = apply_act(a.w, act);
        *reinterpret_cast<float4*>(ob + (size_t)t * out_ld) = a;
      }
    }
  }
}

inline cudaError_t launch_dwconv3x3(const float* in, int in_ld, int in_off, float* out, int out_ld,
                                    int out_off, const float* w, const float* b, int batch, int Hin,
                                    int Win, int C4, int stride, int act, cudaStream_t st) {
  const int Ho = (Hin - 1) / stride + 1, Wo = (Win - 1) / stride + 1;
  int64_t total = (int64_t)batch * Ho * ((Wo + kDwTX - 1) / kDwTX) * (C4 / 4);
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t r = launch_pdl(stride == 1 ? dwconv3x3_kernel<1> : dwconv3x3_kernel<2>, dim3((unsigned)blocks), dim3(256),
                             0, st, in, in_ld, in_off, out, out_ld, out_off, w, b, batch, Hin, Win, C4, act);
  YNB_COUNT_LAUNCH();
  return r;
}

// =====================================================================================
// Pass-through half of a stride-1 ShuffleNetV2 unit: out[m, slot(2i)] = in[m, i], i < half.
// Together with the branch-2 epilogue (which writes slot(2i+1)) this IS
// torch.cat + channel_shuffle (backbone/shufflenetv2.py:70-76, 14-28): no shuffle copy.
// =====================================================================================
__global__ void __launch_bounds__(256)
interleave_copy_kernel(const float* __restrict__ in, int in_ld, float* __restrict__ out, int out_ld,
                       ChanMap omap, int64_t pixels, int half) {
  const int64_t total = pixels * half;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % half);
    int64_t m = i / half;
    out[m * out_ld + omap.slot(2 * c)] = __ldg(in + m * in_ld + c);
  }
}

inline cudaError_t launch_interleave_copy(const float* in, int in_ld, float* out, int out_ld,
                                          ChanMap omap, int64_t pixels, int half, cudaStream_t st) {
  int64_t blocks = (pixels * half + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  interleave_copy_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, in_ld, out, out_ld, omap, pixels, half);
  YNB_COUNT_LAUNCH();
  return cudaGetLastError();
}

// =====================================================================================
// Tap / parity hook: internal NHWC (with channel map) -> canonical NCHW.
// Tiled transpose through shared memory: coalesced on both sides.
// =====================================================================================
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ in, int in_ld, ChanMap imap, float* __restrict__ out,
                    int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int p = p0 + r, c = c0 + tx;
    float v = 0.0f;
    if (p < HW && c < C) v = in[((size_t)b * HW + p) * in_ld + imap.slot(c)];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r, p = p0 + tx;
    if (p < HW && c < C) out[((size_t)b * C + c) * HW + p] = tile[tx][r];
  }
}

inline cudaError_t launch_nhwc_to_nchw(const float* in, int in_ld, ChanMap imap, float* out,
                                       int batch, int C, int HW, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, (C + 31) / 32, batch);
  nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(in, in_ld, imap, out, C, HW);
  YNB_COUNT_LAUNCH();
  return cudaGetLastError();
}

// =====================================================================================
// FPN / PAN merge: out = a + nearest_resample(a2)   (models/yolo_nano.py:291-296;
// F.interpolate default mode 'nearest': up x2 replicates pixels, x0.5 takes [::2, ::2]).
// mode 1: a2 is (H/2 x W/2), mode 2: a2 is (2H x 2W).  All tensors share `ld` = 96.
// =====================================================================================
__global__ void __launch_bounds__(256)
resample_add_kernel(const float* __restrict__ a, const float* __restrict__ a2, float* __restrict__ out,
                    int batch, int H, int W, int ld, int mode) {
  const int groups = ld >> 2;
  const int64_t total = (int64_t)batch * H * W * groups;
  const int H2 = mode == 1 ? H >> 1 : H << 1, W2 = mode == 1 ? W >> 1 : W << 1;
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    int64_t p = i / groups;
    int x = (int)(p % W);
    int y = (int)((p / W) % H);
    int b = (int)(p / ((int64_t)W * H));
    int y2 = mode == 1 ? y >> 1 : y << 1, x2 = mode == 1 ? x >> 1 : x << 1;
    float4 u = __ldg(reinterpret_cast<const float4*>(a + p * ld) + g);
    float4 v = __ldg(reinterpret_cast<const float4*>(a2 + (((size_t)b * H2 + y2) * W2 + x2) * ld) + g);
    u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
    reinterpret_cast<float4*>(out + p * ld)[g] = u;
  }
}

inline cudaError_t launch_resample_add(const float* a, const float* a2, float* out, int batch, int H, int W,
                                       int ld, int mode, cudaStream_t st) {
  int64_t total = (int64_t)batch * H * W * (ld / 4);
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t r = launch_pdl(resample_add_kernel, dim3((unsigned)blocks), dim3(256), 0, st, a, a2, out, batch, H, W, ld, mode);
  YNB_COUNT_LAUNCH();
  return r;
}

}  // namespace ynb
