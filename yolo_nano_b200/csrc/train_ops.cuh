// Training branch at the head boundary: target assignment, the four
// losses with their gradient w.r.t. the raw head maps, the SGD update, and the backward kernels of
// the depthwise / pointwise convolutions.  All HBM-bound streaming kernels: coalesced row access,
// grids sized in multiples of the SM count, deterministic two-stage reductions (no float atomics).
#pragma once
#include "common.cuh"

namespace ynb {

constexpr int kTrainMaxAnchors = 8;
constexpr int kLossBlocks = kNumSMs * 8;      // persistent grid of the loss kernel
constexpr int kLossThreads = 256;

struct TrainLossParams {
  const float* raw[3];     // NHWC [B, HW_l, ld], channel map obj a | cls A + a*C + c | box A(1+C) + 4a + k
  float* grad[3];          // same layout: d(conf + cls + bbox + iou loss) / d raw
  const float* target;     // [B, N, 11]  (tools.py:97-216)
  double* partials;        // [gridDim.x][4]
  int ld, B, A, C, S;
  int grid[3], stride[3];
  int cells[3];            // HW_l
  int cell_off[3];         // prefix of cells
  int cells_total;         // sum HW_l
  float anchors[3][kTrainMaxAnchors][2];
};

__device__ __forceinline__ float sigmoid_precise(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per (image, cell): the cell's A*(1+C+4) logits are one contiguous row (1020 B at
// A=3, C=80) read and written with lane-strided accesses.  Class logits are only read for
// positive anchors (their gradient is zero everywhere else), which is what keeps the kernel at
// "write the gradient map once" traffic.
//   models/yolo_nano.py:333-358: decode (no clamp) -> iou_score -> gt_conf = iou (detached)
//   tools.py:12-34   MSEWithLogitsLoss: 5*pos*(sigmoid(l) - iou)^2 + neg*sigmoid(l)^2
//   tools.py:236-276 CE on positives, BCE-with-logits (txty) and MSE (twth) weighted by
//                    gt_box_scale_weight*mask, SmoothL1(iou, mask) over ALL anchors; each sum / B.
__global__ void __launch_bounds__(kLossThreads) train_loss_kernel(const TrainLossParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const long long warps_total = (long long)gridDim.x * (kLossThreads / 32);
  const long long rows = (long long)p.B * p.cells_total;
  const int A = p.A, C = p.C;
  const int used = A * (1 + C + 4);
  const float invB = 1.0f / (float)p.B;
  const float fS = (float)p.S;
  const long long N = (long long)p.cells_total * A;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};    // lane 0 only carries the scalar losses

  for (long long row = (long long)blockIdx.x * (kLossThreads / 32) + warp_in_block; row < rows; row += warps_total) {
    const int b = (int)(row / p.cells_total);
    const int r = (int)(row - (long long)b * p.cells_total);
    const int lvl = r >= p.cell_off[2] ? 2 : (r >= p.cell_off[1] ? 1 : 0);
    const int cell = r - p.cell_off[lvl];
    const int gy = cell / p.grid[lvl], gx = cell - gy * p.grid[lvl];
    const float fs = (float)p.stride[lvl];
    const float* __restrict__ raw = p.raw[lvl] + ((long long)b * p.cells[lvl] + cell) * p.ld;
    float* __restrict__ g = p.grad[lvl] + ((long long)b * p.cells[lvl] + cell) * p.ld;
    const float* __restrict__ trow = p.target + ((long long)b * N + (long long)p.cell_off[lvl] * A + (long long)cell * A) * 11;
    for (int c = used + lane; c < p.ld; c += 32) g[c] = 0.0f;     // padding channels of the map

    for (int a = 0; a < A; ++a) {
      const float* __restrict__ t = trow + a * 11;
      const float obj = t[0];
      const float mask = obj > 0.0f ? 1.0f : 0.0f;
      const float pos = obj == 1.0f ? 1.0f : 0.0f;
      const float neg = obj == 0.0f ? 1.0f : 0.0f;
      const float lconf = raw[a];
      const int bo = A * (1 + C) + 4 * a;
      const float tx = raw[bo], ty = raw[bo + 1], tw = raw[bo + 2], th = raw[bo + 3];
      // decode_boxes / input_size (models/yolo_nano.py:120-156,337)
      const float sx = sigmoid_precise(tx), sy = sigmoid_precise(ty);
      const float ew = expf(tw), eh = expf(th);
      const float aw = p.anchors[lvl][a][0], ah = p.anchors[lvl][a][1];
      const float cx = __fmul_rn(__fadd_rn(sx, (float)gx), fs), cy = __fmul_rn(__fadd_rn(sy, (float)gy), fs);
      const float bw = __fmul_rn(ew, aw), bh = __fmul_rn(eh, ah);
      const float x1 = __fsub_rn(cx, bw * 0.5f) / fS, y1 = __fsub_rn(cy, bh * 0.5f) / fS;
      const float x2 = __fadd_rn(cx, bw * 0.5f) / fS, y2 = __fadd_rn(cy, bh * 0.5f) / fS;
      // iou_score (tools.py:219-233)
      const float q1x = t[7], q1y = t[8], q2x = t[9], q2y = t[10];
      const float tlx = fmaxf(x1, q1x), tly = fmaxf(y1, q1y), brx = fminf(x2, q2x), bry = fminf(y2, q2y);
      const float wa = __fsub_rn(x2, x1), ha = __fsub_rn(y2, y1);
      const float area_a = __fmul_rn(wa, ha);
      const float area_b = __fmul_rn(__fsub_rn(q2x, q1x), __fsub_rn(q2y, q1y));
      const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
      const float iw = __fsub_rn(brx, tlx), ih = __fsub_rn(bry, tly);
      const float area_i = __fmul_rn(__fmul_rn(iw, ih), en);
      const float uni = __fsub_rn(__fadd_rn(area_a, area_b), area_i);
      const float iou = area_i / uni;

      // objectness (tools.py:12-34); the label is the detached IoU
      const float pc = sigmoid_precise(lconf);
      const float dpi = pc - iou;
      const float l_conf = 5.0f * (pos * dpi * dpi) + neg * pc * pc;
      const float g_conf = (10.0f * pos * dpi + 2.0f * neg * pc) * pc * (1.0f - pc) * invB;
      // iou loss: SmoothL1(iou, mask), beta = 1 (tools.py:273)
      const float d = iou - mask, ad = fabsf(d);
      const float l_iou = ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
      const float g_iou = (ad < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f)) * invB;
      // ... back through iou = I / (Aa + Ab - I)
      const float inv_u = 1.0f / uni;
      const float g_I = g_iou * (inv_u + area_i * inv_u * inv_u);
      const float g_Aa = -g_iou * area_i * inv_u * inv_u;
      const float g_iw = g_I * ih * en, g_ih = g_I * iw * en;
      // torch.max / torch.min of two tensors split the gradient evenly on ties
      const float sx2 = x2 < q2x ? 1.0f : (x2 == q2x ? 0.5f : 0.0f), sx1 = x1 > q1x ? 1.0f : (x1 == q1x ? 0.5f : 0.0f);
      const float sy2 = y2 < q2y ? 1.0f : (y2 == q2y ? 0.5f : 0.0f), sy1 = y1 > q1y ? 1.0f : (y1 == q1y ? 0.5f : 0.0f);
      const float g_x2 = (g_iw * sx2 + g_Aa * ha) / fS, g_x1 = (-g_iw * sx1 - g_Aa * ha) / fS;
      const float g_y2 = (g_ih * sy2 + g_Aa * wa) / fS, g_y1 = (-g_ih * sy1 - g_Aa * wa) / fS;
      float g_tx = (g_x1 + g_x2) * fs * sx * (1.0f - sx);
      float g_ty = (g_y1 + g_y2) * fs * sy * (1.0f - sy);
      float g_tw = (g_x2 - g_x1) * 0.5f * bw;
      float g_th = (g_y2 - g_y1) * 0.5f * bh;
      // box loss (tools.py:267-270)
      const float wm = t[6] * mask;
      float l_box = 0.0f, l_cls = 0.0f;
      if (mask > 0.0f) {
        const float ttx = t[2], tty = t[3], ttw = t[4], tth = t[5];
        const float bce = (fmaxf(tx, 0.0f) - tx * ttx + log1pf(expf(-fabsf(tx)))) +
                          (fmaxf(ty, 0.0f) - ty * tty + log1pf(expf(-fabsf(ty))));
        const float mse = (tw - ttw) * (tw - ttw) + (th - tth) * (th - tth);
        l_box = bce * wm + mse * wm;
        g_tx += wm * (sx - ttx) * invB;
        g_ty += wm * (sy - tty) * invB;
        g_tw += wm * 2.0f * (tw - ttw) * invB;
        g_th += wm * 2.0f * (th - tth) * invB;
      }
      // class loss: cross entropy on positives (tools.py:264)
      const int co = A + a * C;
      if (mask > 0.0f) {
        const int gt = (int)t[1];
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, raw[co + c]);
        m = warp_max(m);
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s += expf(raw[co + c] - m);
        s = warp_sum(s);
        l_cls = logf(s) + m - raw[co + gt];
        const float inv_s = 1.0f / s;
        for (int c = lane; c < C; c += 32)
          g[co + c] = (expf(raw[co + c] - m) * inv_s - (c == gt ? 1.0f : 0.0f)) * invB;
      } else {
        for (int c = lane; c < C; c += 32) g[co + c] = 0.0f;
      }
      if (lane == 0) {
        g[a] = g_conf;
        acc[0] += l_conf; acc[1] += l_cls; acc[2] += l_box; acc[3] += l_iou;
      }
      if (lane < 4) g[bo + lane] = lane == 0 ? g_tx : (lane == 1 ? g_ty : (lane == 2 ? g_tw : g_th));
    }
  }

  __shared__ double part[kLossThreads / 32][4];
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) part[warp_in_block][k] = (double)acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) s += part[w][threadIdx.x];
    p.partials[(long long)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

// Second stage: fixed-order sum of the block partials, / batch_size.
__global__ void train_loss_finalize_kernel(const double* __restrict__ partials, int blocks, int B, float* __restrict__ losses) {
  __shared__ double sh[4][32];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;     // 128 threads: one warp per loss
  double s = 0.0;
  for (int i = lane; i < blocks; i += 32) s += partials[(long long)i * 4 + k];
  sh[k][lane] = s;
  __syncthreads();
  if (lane == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[k][i];
    losses[k] = (float)(t / (double)B);
  }
}

inline cudaError_t launch_train_loss(TrainLossParams p, float* losses, cudaStream_t st) {
  const long long rows = (long long)p.B * p.cells_total;
  const long long need = (rows + kLossThreads / 32 - 1) / (kLossThreads / 32);
  const int blocks = (int)(need < kLossBlocks ? need : kLossBlocks);
  train_loss_kernel<<<blocks, kLossThreads, 0, st>>>(p);
  YNB_COUNT_LAUNCH();
  train_loss_finalize_kernel<<<1, 128, 0, st>>>(p.partials, blocks, p.B, losses);
  YNB_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ---- torch.optim.SGD(momentum, weight_decay), one step over a flat parameter vector ------------
// train.py:167-171,230.  d = g + wd*p;  buf = first ? d : m*buf + d;  p -= lr*buf.
// Same operation order as torch (add with alpha = one fused multiply-add), so the update is
// bit-identical to the CPU optimiser.  16-byte accesses, 3 reads + 2 writes per element.
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ buf, long long n, float lr, float momentum,
                                                       float wd, int first, float grad_scale) {
  const long long n4 = n >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  auto upd = [&](float pv, float gv, float bv, float& pn, float& bn) {
    const float gs = grad_scale == 1.0f ? gv : __fmul_rn(gv, grad_scale);
    const float d = wd != 0.0f ? __fmaf_rn(pv, wd, gs) : gs;
    bn = first ? d : __fadd_rn(__fmul_rn(bv, momentum), d);
    pn = __fmaf_rn(bn, -lr, pv);
  };
  for (long long i = tid; i < n4; i += nth) {
    const float4 pv = reinterpret_cast<const float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(buf)[i];
    float4 pn, bn;
    upd(pv.x, gv.x, bv.x, pn.x, bn.x); upd(pv.y, gv.y, bv.y, pn.y, bn.y);
    upd(pv.z, gv.z, bv.z, pn.z, bn.z); upd(pv.w, gv.w, bv.w, pn.w, bn.w);
    reinterpret_cast<float4*>(p)[i] = pn;
    reinterpret_cast<float4*>(buf)[i] = bn;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    float pn, bn;
    upd(p[i], g[i], first ? 0.0f : buf[i], pn, bn);
    p[i] = pn; buf[i] = bn;
  }
}

inline cudaError_t launch_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                                   float wd, int first, float grad_scale, cudaStream_t st) {
  long long need = ((n >> 2) + 255) / 256;
  if (need < 1) need = 1;
  const int blocks = (int)(need < kNumSMs * 8 ? need : kNumSMs * 8);
  sgd_step_kernel<<<blocks, 256, 0, st>>>(p, g, buf, n, lr, momentum, wd, first, grad_scale);
  YNB_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ---- tools.multi_gt_creator (tools.py:97-216) on the device ----------------------------------------
// labels [B, L, 5] float32 = xmin, ymin, xmax, ymax (normalised), class; counts [B].
// One thread per image walks its labels IN ORDER (later labels overwrite earlier ones on the same
// cell / anchor, as the reference's Python loop does); float64 arithmetic in the reference's
// operation order (no FMA contraction), cast to float32 on store.  target must be zeroed before.
struct TargetParams {
  const float* labels;
  const int* counts;
  float* target;           // [B, N, 11]
  int B, L, A, S;
  int grid[3], stride[3];
  long long row_off[3];    // first anchor row of each level
  long long N;
  double anchors[3 * kTrainMaxAnchors][2];
};

__global__ void build_targets_kernel(const TargetParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int n = p.counts ? min(p.counts[b], p.L) : p.L;
  const double w = (double)p.S, h = (double)p.S;
  const int na = 3 * p.A;
  float* __restrict__ tg = p.target + (long long)b * p.N * 11;
  for (int i = 0; i < n; ++i) {
    const float* lab = p.labels + ((long long)b * p.L + i) * 5;
    const double xmin = lab[0], ymin = lab[1], xmax = lab[2], ymax = lab[3];
    const int cls = (int)lab[4];
    const double cx = __dmul_rn(__dadd_rn(xmax, xmin) / 2.0, w), cy = __dmul_rn(__dadd_rn(ymax, ymin) / 2.0, h);
    const double bw = __dmul_rn(__dsub_rn(xmax, xmin), w), bh = __dmul_rn(__dsub_rn(ymax, ymin), h);
    if (bw < 1.0 || bh < 1.0) continue;                       // "dirty data" (:120-122)
    // compute_iou (tools.py:37-77) of [0,0,aw,ah] against [0,0,bw,bh]
    double best_iou = 0.0;
    int best = 0;
    unsigned above = 0;
    for (int k = 0; k < na; ++k) {
      const double aw = p.anchors[k][0], ah = p.anchors[k][1];
      const double iw = __dsub_rn(fmin(0.0 + bw / 2.0, 0.0 + aw / 2.0), fmax(0.0 - bw / 2.0, 0.0 - aw / 2.0));
      const double ih = __dsub_rn(fmin(0.0 + bh / 2.0, 0.0 + ah / 2.0), fmax(0.0 - bh / 2.0, 0.0 - ah / 2.0));
      const double si = __dmul_rn(ih, iw);
      const double u = __dadd_rn(__dsub_rn(__dadd_rn(__dmul_rn(bw, bh), __dmul_rn(aw, ah)), si), 1e-20);
      const double iou = si / u;
      if (k == 0 || iou > best_iou) { best_iou = iou; best = k; }      // np.argmax: first maximum
      if (iou > 0.5) above |= 1u << k;                                  // IGNORE_THRESH (data/config.py:3)
    }
    for (int k = 0; k < na; ++k) {
      if (!(k == best || ((above >> k) & 1u))) continue;
      const int lvl = k / p.A, a = k - lvl * p.A;
      const double s = (double)p.stride[lvl];
      const double cxs = cx / s, cys = cy / s;
      const int gx = (int)cxs, gy = (int)cys;
      if (gx >= p.grid[lvl] || gy >= p.grid[lvl]) continue;   // :147,:186 (the reference raises for ignored ones)
      float* t = tg + (p.row_off[lvl] + ((long long)gy * p.grid[lvl] + gx) * p.A + a) * 11;
      if (k == best) {
        t[0] = 1.0f;
        t[1] = (float)cls;
        t[2] = (float)__dsub_rn(cxs, (double)gx);
        t[3] = (float)__dsub_rn(cys, (double)gy);
        t[4] = (float)log(bw / p.anchors[k][0]);
        t[5] = (float)log(bh / p.anchors[k][1]);
        t[6] = (float)__dsub_rn(2.0, __dmul_rn(bw / w, bh / h));
        t[7] = (float)xmin; t[8] = (float)ymin; t[9] = (float)xmax; t[10] = (float)ymax;
      } else {
        t[0] = -1.0f;
        t[6] = -1.0f;
      }
    }
  }
}

inline cudaError_t launch_build_targets(const TargetParams& p, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(p.target, 0, (size_t)p.B * p.N * 11 * sizeof(float), st);
  if (e != cudaSuccess) return e;
  build_targets_kernel<<<(p.B + 31) / 32, 32, 0, st>>>(p);
  YNB_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ---- backward of the depthwise 3x3 convolution (config 5: "backward of the dw/pw conv kernels") -----
// Forward (ynb_dwconv3x3): out[b,y,x,c] = bias[c] + sum_t w[t][c] * in[b, y*s+dy-1, x*s+dx-1, c].
// NHWC, channels innermost: a thread owns 4 channels of one pixel (16-byte accesses).
//   dIn[b,yi,xi,c] = sum_t w[t][c] * dOut[b, (yi+1-dy)/s, (xi+1-dx)/s, c]   (where divisible and inside)
__global__ void __launch_bounds__(256) dwconv3x3_bwd_data_kernel(const float* __restrict__ dout, int do_ld, int do_off,
                                                                 float* __restrict__ din, int di_ld, int di_off,
                                                                 const float* __restrict__ w, int B, int h_in, int w_in,
                                                                 int C, int stride) {
  const int c4n = (C + 3) >> 2;
  const int h_out = (h_in - 1) / stride + 1, w_out = (w_in - 1) / stride + 1;
  const long long total = (long long)B * h_in * w_in * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    long long pix = i / c4n;
    const int xi = (int)(pix % w_in);
    pix /= w_in;
    const int yi = (int)(pix % h_in);
    const int b = (int)(pix / h_in);
    const int c = c4 * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ty = yi + 1 - dy;
      if (ty < 0 || ty % stride) continue;
      const int yo = ty / stride;
      if (yo >= h_out) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int txx = xi + 1 - dx;
        if (txx < 0 || txx % stride) continue;
        const int xo = txx / stride;
        if (xo >= w_out) continue;
        const float* dp = dout + (((long long)b * h_out + yo) * w_out + xo) * do_ld + do_off + c;
        const float* wp = w + (dy * 3 + dx) * C + c;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < C) acc[k] = fmaf(wp[k], dp[k], acc[k]);
      }
    }
    float* op = din + (((long long)b * h_in + yi) * w_in + xi) * di_ld + di_off + c;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c + k < C) op[k] = acc[k];
  }
}

//   dW[t][c] = sum_{b,y,x} dOut[b,y,x,c] * in[b, y*s+dy-1, x*s+dx-1, c];  dBias[c] = sum dOut[b,y,x,c]
// Stage 1: block (row-chunk j, channel group) accumulates 10 sums per channel over its output rows in
// registers (thread = channel x pixel-slice), reduces across the pixel slices in shared memory and
// writes partial[j][10][C].  Stage 2 (reduce_partials_kernel) sums the chunks in fixed order.
constexpr int kDwBwdChan = 32;     // channels per block (x)
constexpr int kDwBwdSlices = 8;    // pixel slices per block (y)
__global__ void __launch_bounds__(kDwBwdChan * kDwBwdSlices)
dwconv3x3_bwd_weight_kernel(const float* __restrict__ dout, int do_ld, int do_off, const float* __restrict__ in, int in_ld,
                            int in_off, float* __restrict__ partial, int B, int h_in, int w_in, int C, int stride,
                            int rows_per_chunk) {
  const int h_out = (h_in - 1) / stride + 1, w_out = (w_in - 1) / stride + 1;
  const int c = blockIdx.y * kDwBwdChan + threadIdx.x;
  const long long rows = (long long)B * h_out;                 // (b, yo) rows
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.0f;
  if (c < C) {
    for (long long r = r0; r < r1; ++r) {
      const int b = (int)(r / h_out), yo = (int)(r % h_out);
      for (int xo = threadIdx.y; xo < w_out; xo += kDwBwdSlices) {
        const float d = dout[(((long long)b * h_out + yo) * w_out + xo) * do_ld + do_off + c];
        acc[9] += d;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yi = yo * stride + dy - 1;
          if (yi < 0 || yi >= h_in) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int xi = xo * stride + dx - 1;
            if (xi < 0 || xi >= w_in) continue;
            acc[dy * 3 + dx] = fmaf(d, in[(((long long)b * h_in + yi) * w_in + xi) * in_ld + in_off + c], acc[dy * 3 + dx]);
          }
        }
      }
    }
  }
  __shared__ float sh[kDwBwdSlices][10][kDwBwdChan];
#pragma unroll
  for (int k = 0; k < 10; ++k) sh[threadIdx.y][k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int k = threadIdx.y; k < 10; k += kDwBwdSlices) {
    float s = 0.0f;
    for (int j = 0; j < kDwBwdSlices; ++j) s += sh[j][k][threadIdx.x];
    if (c < C) partial[((long long)blockIdx.x * 10 + k) * C + c] = s;
  }
}

// out[i] = sum_j partial[j][i], j in fixed order (deterministic second stage of the weight gradients).
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int chunks, long long elems, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int j = 0; j < chunks; ++j) s += partial[(long long)j * elems + i];
    out[i] = s;
  }
}

inline int dw_bwd_chunks(int B, int h_out) {
  const long long rows = (long long)B * h_out;
  long long chunks = kNumSMs * 2;
  if (chunks > rows) chunks = rows;
  return (int)chunks;
}

// ---- backward of the pointwise conv w.r.t. its weights -----------------------------------------------
// dW[n][k] = sum_m dOut[m, n] * in[m, k];  dBias[n] = sum_m dOut[m, n].   (dIn = dOut . W is the forward
// GEMM with the transposed weight matrix: ynb_pwconv / ynb_pwconv_tc.)
// A reduction over M (10^4 ... 10^6 pixels) into a small N x K matrix: block (chunk of M, 64 x 64 tile of
// (n, k)) stages 32-pixel slabs of dOut and in through shared memory, each thread owns a 4 x 4
// register tile; partial[chunk][N][K] is summed in fixed order by reduce_partials_kernel.
constexpr int kPwBwdTile = 64;
constexpr int kPwBwdSlab = 32;
__global__ void __launch_bounds__(256) pwconv_bwd_weight_kernel(const float* __restrict__ dout, int do_ld, int do_off,
                                                                const float* __restrict__ in, int in_ld, int in_off,
                                                                float* __restrict__ partial_w, float* __restrict__ partial_b,
                                                                long long M, int K, int N, long long m_per_chunk) {
  __shared__ float sd[kPwBwdSlab][kPwBwdTile + 4];
  __shared__ float sx[kPwBwdSlab][kPwBwdTile + 4];
  const int n0 = blockIdx.y * kPwBwdTile, k0 = blockIdx.z * kPwBwdTile;
  const long long m0 = (long long)blockIdx.x * m_per_chunk;
  const long long m1 = m0 + m_per_chunk < M ? m0 + m_per_chunk : M;
  const int tn = (threadIdx.x >> 4) * 4, tk = (threadIdx.x & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bsum = 0.0f;                                        // threads 0..63 (k-tile 0 only): bias gradient
  for (long long m = m0; m < m1; m += kPwBwdSlab) {
    for (int i = threadIdx.x; i < kPwBwdSlab * kPwBwdTile; i += 256) {
      const int r = i / kPwBwdTile, cidx = i % kPwBwdTile;
      const long long mm = m + r;
      sd[r][cidx] = (mm < m1 && n0 + cidx < N) ? dout[mm * do_ld + do_off + n0 + cidx] : 0.0f;
      sx[r][cidx] = (mm < m1 && k0 + cidx < K) ? in[mm * in_ld + in_off + k0 + cidx] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kPwBwdSlab; ++r) {
      const float4 dv = *reinterpret_cast<const float4*>(&sd[r][tn]);
      const float4 xv = *reinterpret_cast<const float4*>(&sx[r][tk]);
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], xx[j], acc[i][j]);
    }
    if (blockIdx.z == 0 && threadIdx.x < kPwBwdTile) {
      for (int r = 0; r < kPwBwdSlab; ++r) bsum += sd[r][threadIdx.x];
    }
    __syncthreads();
  }
  float* pw = partial_w + (long long)blockIdx.x * N * K;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + tn + i < N && k0 + tk + j < K) pw[(long long)(n0 + tn + i) * K + k0 + tk + j] = acc[i][j];
  if (blockIdx.z == 0 && threadIdx.x < kPwBwdTile && n0 + threadIdx.x < N)
    partial_b[(long long)blockIdx.x * N + n0 + threadIdx.x] = bsum;
}

inline int pw_bwd_chunks(long long M, int K, int N) {
  const int tiles = ((N + kPwBwdTile - 1) / kPwBwdTile) * ((K + kPwBwdTile - 1) / kPwBwdTile);
  long long chunks = (kNumSMs * 4 + tiles - 1) / tiles;       // ~4 blocks per SM in total
  const long long max_chunks = (M + kPwBwdSlab - 1) / kPwBwdSlab;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

// dOut *= act'(out)  in place is avoided: writes dPre[m, c] = dOut[m, c] * (out > 0 ? 1 : slope), slope = 0
// (ReLU) or 0.1 (LeakyReLU), for channel ranges given as (ld, off).  `out` is the forward OUTPUT (its sign
// equals the sign of the pre-activation for both activations).
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dout, int do_ld, int do_off,
                                                      const float* __restrict__ out, int o_ld, int o_off,
                                                      float* __restrict__ dpre, int dp_ld, int dp_off, long long M, int C,
                                                      float slope) {
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / C;
    const int c = (int)(i - m * C);
    const float o = out[m * o_ld + o_off + c];
    dpre[m * dp_ld + dp_off + c] = dout[m * do_ld + do_off + c] * (o > 0.0f ? 1.0f : slope);
  }
}

}  // namespace ynb
