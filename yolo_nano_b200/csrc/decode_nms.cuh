// Decode of the raw head outputs and batched per-image / per-class greedy NMS.
// Replaces models/yolo_nano.py:303-330 (re-layout), :120-156 (box decode), :362-367
// (sigmoid / softmax / clamp), :245-279 (postprocess) and :159-242 (nms, diou_nms), which
// the reference runs as ~15 ATen kernels, two D2H copies and Python/NumPy loops.
//
// Exactness contract: every IoU expression below is evaluated in float32 in
// the reference's operand order with IEEE add/mul/div/sqrt (the intrinsics stop the
// compiler from contracting into FMA), so for identical candidate boxes the keep-set is
// bit-identical to the NumPy loop.  Visiting order inside a class: score descending, equal
// scores by ascending anchor index (NumPy's argsort is unstable there).
#pragma once
#include "common.cuh"

namespace ynb {

// =====================================================================================
// decode: one thread per (image, cell, anchor)
// raw: NHWC [B, G*G, ld], channel map of models/yolo_nano.py:312-318:
//   ch a            objectness of anchor a
//   ch A + a*C + c  class c of anchor a
//   ch A(1+C)+4a+k  t_x, t_y, t_w, t_h of anchor a
// =====================================================================================
struct DecodeParams {
  const float* raw;
  int ld;
  float* boxes;    // [B,N,4]
  float* scores;   // [B,N]
  int32_t* cls;    // [B,N]
  int batch, G, A, C;
  float stride, input_size;
  float anchor_w[4], anchor_h[4];
  int64_t N, level_off;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// A block stages kDecCells consecutive cells (whole NHWC rows, contiguous in memory) in shared
// memory with 16-byte coalesced loads; then one thread per (cell, anchor) reads its 1+C+4
// values from there.  The raw head map is read exactly once, at full line efficiency.
constexpr int kDecCells = 32;

__global__ void __launch_bounds__(kDecCells * 3)
decode_level_kernel(DecodeParams p) {
  extern __shared__ float s_raw[];               // [kDecCells][pitch]
  const int pitch = p.ld + 1;                    // odd pitch: cells land in different banks
  const int64_t cells_total = (int64_t)p.batch * p.G * p.G;
  const int64_t cell0 = (int64_t)blockIdx.x * kDecCells;
  const int ncell = (int)min((int64_t)kDecCells, cells_total - cell0);
  pdl_trigger();
  pdl_wait();
  {
    const float4* src = reinterpret_cast<const float4*>(p.raw + cell0 * p.ld);
    const int ld4 = p.ld >> 2;
    for (int i = threadIdx.x; i < ncell * ld4; i += blockDim.x) {
      float4 v = __ldg(src + i);
      int cc = i / ld4, k = (i - cc * ld4) << 2;
      float* d = s_raw + cc * pitch + k;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  }
  __syncthreads();
  const int cc = threadIdx.x / p.A, a = threadIdx.x - cc * p.A;
  if (cc >= ncell) return;
  const int64_t cell = cell0 + cc;                // b*G*G + y*G + x
  const int gx = (int)(cell % p.G);
  const int gy = (int)((cell / p.G) % p.G);
  const int b = (int)(cell / ((int64_t)p.G * p.G));
  const float* r = s_raw + cc * pitch;

  const float obj = sigmoidf_(r[a]);
  // softmax over classes (torch.softmax, dim = classes) times objectness, arg-max class.  The
  // reference takes np.argmax over the products (models/yolo_nano.py:253); products are monotone
  // in the logit, so the first maximal logit is that class, and its softmax is exp(0)/sum.
  const float* cl = r + p.A + a * p.C;
  float mx = -INFINITY;
  int best_c = 0;
  for (int c = 0; c < p.C; ++c)
    if (cl[c] > mx) { mx = cl[c]; best_c = c; }
  float sum = 0.0f;
  for (int c = 0; c < p.C; ++c) sum += softmax_exp(cl[c] - mx);
  const float best = class_score(sum, obj);
  // box (models/yolo_nano.py:129-134, 150-154, 366)
  const float* t = r + p.A * (1 + p.C) + 4 * a;
  float tx = t[0], tyv = t[1], tw = t[2], th = t[3];
  const float aw = a == 0 ? p.anchor_w[0] : (a == 1 ? p.anchor_w[1] : (a == 2 ? p.anchor_w[2] : p.anchor_w[3]));
  const float ah = a == 0 ? p.anchor_h[0] : (a == 1 ? p.anchor_h[1] : (a == 2 ? p.anchor_h[2] : p.anchor_h[3]));
  float cx = __fmul_rn(__fadd_rn(sigmoidf_(tx), (float)gx), p.stride);
  float cy = __fmul_rn(__fadd_rn(sigmoidf_(tyv), (float)gy), p.stride);
  float w = __fmul_rn(expf(tw), aw);
  float h = __fmul_rn(expf(th), ah);
  float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
  float x1 = __fdiv_rn(__fsub_rn(cx, hw), p.input_size);
  float y1 = __fdiv_rn(__fsub_rn(cy, hh), p.input_size);
  float x2 = __fdiv_rn(__fadd_rn(cx, hw), p.input_size);
  float y2 = __fdiv_rn(__fadd_rn(cy, hh), p.input_size);
  float4 box = make_float4(fminf(fmaxf(x1, 0.f), 1.f), fminf(fmaxf(y1, 0.f), 1.f),
                           fminf(fmaxf(x2, 0.f), 1.f), fminf(fmaxf(y2, 0.f), 1.f));
  const int64_t n = p.level_off + ((int64_t)gy * p.G + gx) * p.A + a;
  const int64_t o = (int64_t)b * p.N + n;
  reinterpret_cast<float4*>(p.boxes)[o] = box;
  p.scores[o] = best;
  p.cls[o] = best_c;
}

inline cudaError_t launch_decode_level(const DecodeParams& p, cudaStream_t st) {
  int64_t cells = (int64_t)p.batch * p.G * p.G;
  if (cells <= 0) return cudaSuccess;
  if (p.A != 3 || (p.ld & 3)) return cudaErrorInvalidValue;
  size_t smem = (size_t)kDecCells * (p.ld + 1) * 4;
  cudaError_t r = launch_pdl(decode_level_kernel, dim3((unsigned)((cells + kDecCells - 1) / kDecCells)), dim3(kDecCells * 3), smem, st, p);
  YNB_COUNT_LAUNCH();
  return r;
}

// =====================================================================================
// NMS stage 1: sort keys.  key = class(8) | ~score_bits(32) | anchor(16): ascending key
// order = class ascending, score descending, anchor ascending.  Below-threshold anchors
// get the all-ones key and sink to the end (models/yolo_nano.py:258-262: score >= thresh).
// =====================================================================================
constexpr uint64_t kInvalidKey = ~0ull;

__device__ __forceinline__ uint64_t make_key(int cls, float score, int idx) {
  return ((uint64_t)(uint32_t)cls << 48) | ((uint64_t)(~__float_as_uint(score)) << 16) | (uint64_t)(uint32_t)idx;
}
__device__ __forceinline__ int key_idx(uint64_t k) { return (int)(k & 0xFFFFull); }
__device__ __forceinline__ int key_cls(uint64_t k) { return (int)(k >> 48); }

__global__ void __launch_bounds__(256)
nms_keys_kernel(const float* __restrict__ scores, const int32_t* __restrict__ cls,
                uint64_t* __restrict__ keys, uint8_t* __restrict__ keep, int64_t N, int Npad,
                float conf_thresh) {
  const int b = blockIdx.y;
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) {
    uint64_t k = kInvalidKey;
    if (i < N) {
      float s = scores[(int64_t)b * N + i];
      if (s >= conf_thresh) k = make_key(cls[(int64_t)b * N + i], s, i);   // NaN fails, as in NumPy
      keep[(int64_t)b * N + i] = 0;
    }
    keys[(int64_t)b * Npad + i] = k;
  }
}

// One CTA per image, bitonic network over Npad (power of two) keys.  Keys are unique
// (anchor index in the low bits), so the network's instability is irrelevant.  After the sort
// the CTA also writes the class segment table seg[b][0..C]: class c owns sorted positions
// [seg[c], seg[c+1]).
template <bool IN_SMEM>
__global__ void __launch_bounds__(1024)
nms_sort_kernel(uint64_t* __restrict__ keys_g, int32_t* __restrict__ seg_g, int Npad, int C) {
  extern __shared__ uint64_t s_keys[];
  uint64_t* g = keys_g + (int64_t)blockIdx.x * Npad;
  uint64_t* k = IN_SMEM ? s_keys : g;
  int32_t* seg = seg_g + (int64_t)blockIdx.x * (C + 1);
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (IN_SMEM) {
    for (int i = tid; i < Npad; i += 1024) s_keys[i] = g[i];
    __syncthreads();
  }
  for (int size = 2; size <= Npad; size <<= 1) {
    for (int j = size >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (Npad >> 1); t += 1024) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
        int hi = lo | j;
        bool up = (lo & size) == 0;
        uint64_t a = k[lo], c = k[hi];
        if ((a > c) == up) { k[lo] = c; k[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i <= Npad; i += 1024) {
    int ci = i < Npad ? min(key_cls(k[i]), C) : C;          // virtual sentinel of class C at the end
    int cp = i > 0 ? min(key_cls(k[i - 1]), C) : -1;
    for (int c = cp + 1; c <= ci; ++c) seg[c] = i;
    if (IN_SMEM && i < Npad) g[i] = k[i];
  }
}

// =====================================================================================
// NMS stage 2: one CTA per (image, class) segment of the sorted keys; greedy suppression
// in chunks of 64 candidates:
//   (1) 64x64 suppression bitmask inside the chunk (warp-cooperative),
//   (2) one thread resolves the chunk serially against the mask (64-bit ops),
//   (3) all threads test every later candidate against the boxes kept in this chunk.
// Total pair tests = kept x later, the same work as the reference loop, but parallel.
// =====================================================================================
struct IouOps {
  float thr;
  bool diou;
};

__device__ __forceinline__ float box_area(float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));   // models/yolo_nano.py:166
}

// true when kept box `a` suppresses candidate `b`:  NOT (ovr <= thr)   (:185, NaN suppresses)
//
// The reference decision is RN(inter / union) <= thr.  Almost every pair is far from the
// threshold, so it is decided without the IEEE division: with q = inter/union exact,
//   inter < thr*union*(1-1e-6)  =>  q < thr*(1-8e-7)  =>  RN(q) <= thr      (keep)
//   inter > thr*union*(1+1e-6)  =>  q > thr*(1+8e-7)  =>  RN(q) >  thr      (suppress)
// (the two fp32 roundings of the right-hand side cost < 2.4e-7 relative; the next float
// above thr is at most thr*(1+2.4e-7) away).  Everything else — the 2e-6 band around the
// threshold, zero / negative / non-finite unions — takes the exact path, so the result is
// bit-identical to the NumPy loop.  The division's slow path (tiny numerators of
// non-overlapping pairs, inter ~ 1e-28*h) is what made the exact test expensive.
__device__ __forceinline__ bool suppresses(float4 a, float area_a, float4 b, float area_b, IouOps op) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(1e-28f, __fsub_rn(xx2, xx1));
  float h = fmaxf(1e-28f, __fsub_rn(yy2, yy1));
  float inter = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (!op.diou && uni > 1e-20f && uni < 4.0f) {
    float tu = __fmul_rn(op.thr, uni);
    if (inter < __fmul_rn(tu, 0.999999f)) return false;
    if (inter > __fmul_rn(tu, 1.000001f)) return true;
  }
  float ovr = __fdiv_rn(inter, uni);
  if (op.diou) {   // models/yolo_nano.py:216-237
    float cw = __fsub_rn(fmaxf(fmaxf(a.x, a.z), fmaxf(b.x, b.z)), fminf(fminf(a.x, a.z), fminf(b.x, b.z)));
    float ch = __fsub_rn(fmaxf(fmaxf(a.y, a.w), fmaxf(b.y, b.w)), fminf(fminf(a.y, a.w), fminf(b.y, b.w)));
    float C = __fsqrt_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)));
    float p1x = __fmul_rn(__fadd_rn(a.x, a.z), 0.5f), p1y = __fmul_rn(__fadd_rn(a.y, a.w), 0.5f);
    float p2x = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f), p2y = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
    float dx = __fsub_rn(p2x, p1x), dy = __fsub_rn(p2y, p1y);
    float D = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    float lens = __fdiv_rn(__fmul_rn(D, D), __fadd_rn(__fmul_rn(C, C), 1e-20f));
    ovr = __fsub_rn(ovr, lens);
  }
  return !(ovr <= op.thr);
}

// 8x8 occupancy mask of a box on the unit square (closed cell ranges): two boxes whose masks
// do not intersect are disjoint, so w or h is clamped to 1e-28, ovr <= 1e-16 and the pair
// can never suppress — unless an area is ~0 (0/0 = NaN suppresses, SURVEY hazard 3) or a
// coordinate is not finite; such boxes get the all-ones mask and always take the full test.
__device__ __forceinline__ unsigned long long coarse_mask(float4 b, float area) {
  if (!(area > 1e-12f) || !(b.x >= 0.f && b.y >= 0.f && b.z <= 1.f && b.w <= 1.f)) return ~0ull;
  int cx1 = min(7, (int)(b.x * 8.f)), cx2 = min(7, (int)(b.z * 8.f));
  int cy1 = min(7, (int)(b.y * 8.f)), cy2 = min(7, (int)(b.w * 8.f));
  unsigned long long row = (unsigned long long)((2u << cx2) - (1u << cx1)) & 0xffull;
  unsigned long long m = 0ull;
#pragma unroll
  for (int cy = 0; cy < 8; ++cy)
    if (cy >= cy1 && cy <= cy2) m |= row << (8 * cy);
  return m;
}

constexpr int kNmsChunk = 512;                 // candidates resolved per round (= threads)
constexpr int kNmsThreads = 512;
constexpr int kNmsWpr = kNmsChunk / 64;        // 64-bit words per suppression row
constexpr int kNmsMaxWords = 1024;             // 65536 candidates per segment, 64 per word

struct NmsSmem {
  unsigned long long removed[kNmsMaxWords];          // segment-wide "suppressed" bitmap
  float4 cbox[kNmsChunk];                            // chunk candidates, visiting order
  float carea[kNmsChunk];
  unsigned long long cmask[kNmsChunk];               // coarse 8x8 occupancy masks
  unsigned long long rows[kNmsChunk][kNmsWpr];       // rows[r] bit j: r suppresses j (j > r)
  unsigned long long cellocc[64][kNmsWpr];           // coarse cell -> chunk candidates touching it
  unsigned long long kept[kNmsWpr];                  // chunk candidates kept by this round
};

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// One CTA per (image, class) segment; rounds of 512 candidates:
//   A  load the chunk (boxes, areas, coarse masks)
//   B  coarse-cell inverted index of the chunk (cell -> 512-bit set of candidates touching it)
//   C  suppression rows: candidate r against the later candidates that share a coarse cell
//   D  one warp resolves the chunk greedily, 64 rows at a time, with shuffles on the diagonal
//      64x64 blocks and deferred OR-propagation to the later words
//   E  every later candidate of the segment against the kept boxes that share a cell with it
// The pair tests are exactly those of the reference loop (kept box vs every later survivor),
// minus pairs that cannot overlap.
__global__ void __launch_bounds__(kNmsThreads)
nms_segment_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ seg,
                   const float* __restrict__ boxes, float4* sbox_scratch, uint8_t* __restrict__ keep,
                   int64_t N, int Npad, int C, IouOps op) {
  extern __shared__ __align__(16) unsigned char nms_smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(nms_smem_raw);

  const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  pdl_trigger();
  pdl_wait();
  const int s0 = seg[(int64_t)b * (C + 1) + c];
  const int n = seg[(int64_t)b * (C + 1) + c + 1] - s0;
  if (n <= 0) return;
  const uint64_t* k = keys + (int64_t)b * Npad;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (int64_t)b * N;
  float4* sb = sbox_scratch + (int64_t)b * Npad + s0;
  uint8_t* kp = keep + (int64_t)b * N;
  const bool use_masks = op.thr >= 1e-6f;

  // gather the segment's boxes in visiting order (coalesced from here on)
  for (int j = tid; j < n; j += kNmsThreads) sb[j] = bx[key_idx(k[s0 + j])];
  for (int j = tid; j < (n + 63) / 64 + kNmsWpr; j += kNmsThreads)
    if (j < kNmsMaxWords) S.removed[j] = 0ull;
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += kNmsChunk) {
    const int cn = min(kNmsChunk, n - c0);
    // ---- A: chunk -------------------------------------------------------------------------
    {
      float4 v = tid < cn ? sb[c0 + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
      float a = box_area(v);
      S.cbox[tid] = v;
      S.carea[tid] = a;
      S.cmask[tid] = tid < cn ? (use_masks ? coarse_mask(v, a) : ~0ull) : 0ull;
    }
    __syncthreads();
    // ---- B: inverted index: thread -> (cell, word) ------------------------------------------
    {
      const int cell = tid >> 3, w = tid & 7;
      unsigned long long m = 0ull;
#pragma unroll 8
      for (int q = 0; q < 64; ++q) m |= ((S.cmask[w * 64 + q] >> cell) & 1ull) << q;
      S.cellocc[cell][w] = m;
    }
    __syncthreads();
    // ---- C: suppression row of candidate `tid` --------------------------------------------------
    {
      unsigned long long rw[kNmsWpr];
#pragma unroll
      for (int w = 0; w < kNmsWpr; ++w) rw[w] = 0ull;
      if (tid < cn) {
        const float4 a = S.cbox[tid];
        const float aa = S.carea[tid];
        unsigned long long am = S.cmask[tid];
        unsigned long long cand[kNmsWpr];
        if (am == ~0ull) {
#pragma unroll
          for (int w = 0; w < kNmsWpr; ++w) cand[w] = ~0ull;
        } else {
#pragma unroll
          for (int w = 0; w < kNmsWpr; ++w) cand[w] = 0ull;
          while (am) {
            const int cell = __ffsll((long long)am) - 1;
            am &= am - 1;
#pragma unroll
            for (int w = 0; w < kNmsWpr; ++w) cand[w] |= S.cellocc[cell][w];
          }
        }
        // only later candidates (j > tid) that exist (j < cn).  Degenerate boxes carry the all-ones
        // mask, i.e. they sit in every cell of the index and are met from any row.
#pragma unroll
        for (int w = 0; w < kNmsWpr; ++w) {
          const int base = w * 64;
          unsigned long long later = tid >= base + 63 ? 0ull : (tid < base ? ~0ull : (~0ull << (tid - base + 1)));
          unsigned long long exist = cn >= base + 64 ? ~0ull : (cn <= base ? 0ull : ((1ull << (cn - base)) - 1ull));
          unsigned long long todo = cand[w] & later & exist;
          while (todo) {
            const int q = __ffsll((long long)todo) - 1;
            todo &= todo - 1;
            const int j = base + q;
            if (suppresses(a, aa, S.cbox[j], S.carea[j], op)) rw[w] |= 1ull << q;
          }
        }
      }
#pragma unroll
      for (int w = 0; w < kNmsWpr; ++w) S.rows[tid][w] = rw[w];
    }
    __syncthreads();
    // ---- D: greedy resolve of the chunk by warp 0 ----------------------------------------------
    if (tid < 32) {
      // rem: lane w (< 8) holds the suppressed bits of block w (from earlier rounds, then from
      // kept rows of earlier blocks of this round); pend[w']: this lane's not-yet-reduced
      // contribution to block w' from the kept rows it owns
      unsigned long long rem = lane < kNmsWpr ? S.removed[(c0 >> 6) + lane] : 0ull;
      unsigned long long pend[kNmsWpr];
#pragma unroll
      for (int w = 0; w < kNmsWpr; ++w) pend[w] = 0ull;
      unsigned long long my_kept = 0ull;
#pragma unroll
      for (int w = 0; w < kNmsWpr; ++w) {
        const int base = w * 64;
        // fold the pending contributions for this block (OR-reduce over the warp)
        unsigned long long inc = pend[w];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) inc |= __shfl_xor_sync(0xffffffffu, inc, d);
        unsigned long long cur = shfl64(rem, w) | inc;
        const unsigned long long exist = cn >= base + 64 ? ~0ull : (cn <= base ? 0ull : ((1ull << (cn - base)) - 1ull));
        // diagonal block: lane owns rows base+lane and base+32+lane
        const unsigned long long d0 = S.rows[base + lane][w], d1 = S.rows[base + 32 + lane][w];
        unsigned long long alive = ~cur & exist, kept_bits = 0ull;
        while (alive) {
          const int j = __ffsll((long long)alive) - 1;
          kept_bits |= 1ull << j;
          const unsigned long long dj = shfl64(j < 32 ? d0 : d1, j & 31);
          alive &= ~dj;
          alive &= ~(1ull << j);
        }
        if (lane == w) my_kept = kept_bits;
        // kept rows owned by this lane suppress candidates of the later blocks
        const bool k0 = (kept_bits >> lane) & 1ull, k1 = (kept_bits >> (lane + 32)) & 1ull;
#pragma unroll
        for (int w2 = w + 1; w2 < kNmsWpr; ++w2) {
          if (k0) pend[w2] |= S.rows[base + lane][w2];
          if (k1) pend[w2] |= S.rows[base + 32 + lane][w2];
        }
      }
      if (lane < kNmsWpr) S.kept[lane] = my_kept;
    }
    __syncthreads();
    // ---- E: flags of the kept candidates; kept boxes vs every later candidate ------------------
    if ((S.kept[tid >> 6] >> (tid & 63)) & 1ull) kp[key_idx(k[s0 + c0 + tid])] = 1;
    for (int j = c0 + kNmsChunk + tid; j < n; j += kNmsThreads) {
      if ((S.removed[j >> 6] >> (j & 63)) & 1ull) continue;
      const float4 v = sb[j];
      const float va = box_area(v);
      unsigned long long vm = use_masks ? coarse_mask(v, va) : ~0ull;
      unsigned long long hits[kNmsWpr];
      if (vm == ~0ull) {
#pragma unroll
        for (int w = 0; w < kNmsWpr; ++w) hits[w] = S.kept[w];
      } else {
#pragma unroll
        for (int w = 0; w < kNmsWpr; ++w) hits[w] = 0ull;
        while (vm) {
          const int cell = __ffsll((long long)vm) - 1;
          vm &= vm - 1;
#pragma unroll
          for (int w = 0; w < kNmsWpr; ++w) hits[w] |= S.cellocc[cell][w];
        }
#pragma unroll
        for (int w = 0; w < kNmsWpr; ++w) hits[w] &= S.kept[w];
      }
      bool gone = false;
#pragma unroll
      for (int w = 0; w < kNmsWpr; ++w) {
        unsigned long long h = hits[w];
        while (h && !gone) {
          const int q = __ffsll((long long)h) - 1;
          h &= h - 1;
          const int r = w * 64 + q;
          if (suppresses(S.cbox[r], S.carea[r], v, va, op)) gone = true;
        }
      }
      if (gone) atomicOr(&S.removed[j >> 6], 1ull << (j & 63));
    }
    __syncthreads();
  }
}

// =====================================================================================
// NMS stage 3: per image, compact kept anchors in ascending anchor order
// (models/yolo_nano.py:274-277).
// =====================================================================================
__global__ void __launch_bounds__(1024)
nms_compact_kernel(const uint8_t* __restrict__ keep, const float* __restrict__ boxes,
                   const float* __restrict__ scores, const int32_t* __restrict__ cls,
                   float* __restrict__ out_boxes, float* __restrict__ out_scores,
                   int32_t* __restrict__ out_cls, int32_t* __restrict__ out_counts, int64_t N) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int per = (int)((N + 1023) / 1024);
  const int beg = min((int)N, tid * per), end = min((int)N, beg + per);
  const uint8_t* kp = keep + (int64_t)b * N;
  pdl_trigger();
  pdl_wait();
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += kp[i];
  // block exclusive scan
  int incl = cnt;
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid], wi = w;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (tid >= d) wi += v;
    }
    s_warp[tid] = wi - w;   // exclusive prefix of warp totals
    if (tid == 31) out_counts[b] = wi;
  }
  __syncthreads();
  int pos = s_warp[tid >> 5] + incl - cnt;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (int64_t)b * N;
  float4* ob = reinterpret_cast<float4*>(out_boxes) + (int64_t)b * N;
  for (int i = beg; i < end; ++i) {
    if (kp[i]) {
      ob[pos] = bx[i];
      out_scores[(int64_t)b * N + pos] = scores[(int64_t)b * N + i];
      out_cls[(int64_t)b * N + pos] = cls[(int64_t)b * N + i];
      ++pos;
    }
  }
}

// ---- host side ------------------------------------------------------------------------
inline int nms_npad(int64_t n) {
  int p = 64;
  while (p < n) p <<= 1;
  return p;
}

struct NmsWorkspace {
  uint64_t* keys;     // [B][Npad]
  float4* sbox;       // [B][Npad]
  int32_t* seg;       // [B][257] class segment starts
  uint8_t* keep;      // [B][N]
};
constexpr int kNmsSegStride = 257;
inline int64_t nms_workspace_bytes(int batch, int64_t n) {
  int64_t npad = nms_npad(n);
  return (int64_t)batch * npad * 8 + (int64_t)batch * npad * 16 + round_up64((int64_t)batch * kNmsSegStride * 4, 256) +
         round_up64((int64_t)batch * n, 256) + 512;
}
inline NmsWorkspace nms_carve(void* ws, int batch, int64_t n) {
  int64_t npad = nms_npad(n);
  char* p = reinterpret_cast<char*>(round_up64((int64_t)(uintptr_t)ws, 256));
  NmsWorkspace w;
  w.sbox = reinterpret_cast<float4*>(p); p += (int64_t)batch * npad * 16;
  w.keys = reinterpret_cast<uint64_t*>(p); p += (int64_t)batch * npad * 8;
  w.seg = reinterpret_cast<int32_t*>(p); p += round_up64((int64_t)batch * kNmsSegStride * 4, 256);
  w.keep = reinterpret_cast<uint8_t*>(p);
  return w;
}

constexpr int kSortSmemLimit = 200 * 1024;

inline cudaError_t launch_nms(const float* boxes, const float* scores, const int32_t* cls, int batch,
                              int64_t N, int num_classes, float conf, float thr, int diou,
                              float* out_boxes, float* out_scores, int32_t* out_cls,
                              int32_t* out_counts, uint8_t* keep_out, NmsWorkspace w, cudaStream_t st) {
  const int Npad = nms_npad(N);
  if (N > 65536 || num_classes > 255) return cudaErrorInvalidValue;
  uint8_t* keep = keep_out ? keep_out : w.keep;
  {
    dim3 grid((Npad + 255) / 256, batch);
    cudaError_t r = launch_pdl(nms_keys_kernel, grid, dim3(256), 0, st, scores, cls, w.keys, keep, N, Npad, conf);
    if (r != cudaSuccess) return r;
    YNB_COUNT_LAUNCH();
  }
  {
    size_t smem = (size_t)Npad * 8;
    if (smem <= (size_t)kSortSmemLimit) {
      static bool attr_set = false;
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(nms_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSortSmemLimit);
        if (e != cudaSuccess) return e;
        attr_set = true;
      }
      cudaError_t r = launch_pdl(nms_sort_kernel<true>, dim3(batch), dim3(1024), smem, st, w.keys, w.seg, Npad, num_classes);
      if (r != cudaSuccess) return r;
    } else {
      cudaError_t r = launch_pdl(nms_sort_kernel<false>, dim3(batch), dim3(1024), 0, st, w.keys, w.seg, Npad, num_classes);
      if (r != cudaSuccess) return r;
    }
    YNB_COUNT_LAUNCH();
  }
  {
    dim3 grid(num_classes, batch);
    IouOps op{thr, diou != 0};
    static bool seg_attr = false;
    if (!seg_attr) {
      cudaError_t e2 = cudaFuncSetAttribute(nms_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(NmsSmem));
      if (e2 != cudaSuccess) return e2;
      seg_attr = true;
    }
    cudaError_t r = launch_pdl(nms_segment_kernel, grid, dim3(kNmsThreads), sizeof(NmsSmem), st, w.keys, w.seg, boxes,
                               w.sbox, keep, N, Npad, num_classes, op);
    if (r != cudaSuccess) return r;
    YNB_COUNT_LAUNCH();
  }
  cudaError_t r = launch_pdl(nms_compact_kernel, dim3(batch), dim3(1024), 0, st, (const uint8_t*)keep, boxes, scores, cls,
                             out_boxes, out_scores, out_cls, out_counts, N);
  YNB_COUNT_LAUNCH();
  return r;
}

}  // namespace ynb
