// Anchor-grid NMS: the product path's replacement for the sort + greedy loop of
// models/yolo_nano.py:159-279, exact (same keep set), without a sort and without a serial scan.
//
// Greedy NMS keeps candidate j iff no KEPT candidate i of the same class that precedes j
// (score descending, ties by ascending anchor index) has ovr(i, j) > thresh (or NaN).  That
// recurrence has exactly one solution, so it can be evaluated in any order:
//
//   prep     one CTA per image: per-anchor meta word (class, candidate, irregular), per
//            (level, anchor) size ranges and class-presence masks, list of irregular candidates
//   build    one thread per candidate: its SUPPRESSOR LIST = preceding same-class candidates
//            with ovr > thresh.  Candidates are anchors of a known grid: the decoded centre of
//            anchor i lies in its cell and inside its (clipped) box, and ovr > t bounds both the
//            size ratio (t < w_i/w_j < 1/t) and the centre distance (|dm| < r(t) w_j), so a
//            candidate only looks at a window of cells around itself, in the (level, anchor)
//            combinations whose size range and class mask admit a partner.  Boxes for which the
//            geometric argument does not hold (zero / denormal area, out of the unit square,
//            not touching their own cell — never produced by the decode, but accepted) are
//            "irregular": they test everything and everything tests them.
//   resolve  one CTA per image, fixed-point iteration in shared memory: j is REMOVED once a
//            listed suppressor is KEPT, KEPT once all of them are REMOVED.  The first
//            candidate in order never has an undecided suppressor, so every sweep decides at
//            least one candidate; typical depth is 5-20 sweeps.  Candidates with more than
//            kSupCap suppressors re-run the window search against the current states.
//   compact  nms_compact_kernel (anchor order = the reference's output order, :274-277)
//
// Pair tests are `suppresses()` of decode_nms.cuh, the same IEEE sequence as the reference.
#pragma once
#include "decode_nms.cuh"

namespace ynb {

constexpr int kSupCap = 32;            // listed suppressors per candidate
constexpr uint32_t kMetaIrregular = 0x80000000u;
constexpr uint32_t kMetaClsMask = 0x1ffu;   // class + 1 (0 = not a candidate)

struct GridGeom {
  int G[3];      // cells per side, level 0..2 (stride 8, 16, 32)
  int off[4];    // first anchor index of each level; off[3] = N
  int N;
};

inline GridGeom make_grid_geom(int input_size) {
  GridGeom g{};
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    g.G[l] = input_size / (8 << l);
    g.off[l] = off;
    off += g.G[l] * g.G[l] * 3;
  }
  g.off[3] = off;
  g.N = off;
  return g;
}

struct ComboStats {                        // per (image, level, anchor), regular candidates only
  uint32_t wmin, wmax, hmin, hmax;         // float bits (positive floats order like uints)
  uint32_t cls[8];                         // classes present, bit c
};

struct GridNmsWorkspace {
  ComboStats* stats;   // [B][9]
  uint32_t* meta;      // [B][N]
  int32_t* dcount;     // [B]
  uint16_t* dlist;     // [B][N] irregular candidates
  uint16_t* cnt;       // [B][N] suppressors found (saturating); > kSupCap: list incomplete
  uint16_t* sup;       // [B][N][kSupCap]
  uint8_t* keep;       // [B][N]
};

inline int64_t grid_nms_workspace_bytes(int batch, int64_t n) {
  int64_t b = batch;
  return round_up64(b * 9 * (int64_t)sizeof(ComboStats), 256) + round_up64(b * n * 4, 256) + round_up64(b * 4, 256) +
         round_up64(b * n * 2, 256) + round_up64(b * n * 2, 256) + round_up64(b * n * kSupCap * 2, 256) +
         round_up64(b * n, 256) + 512;
}
inline GridNmsWorkspace grid_nms_carve(void* ws, int batch, int64_t n) {
  int64_t b = batch;
  char* p = reinterpret_cast<char*>(round_up64((int64_t)(uintptr_t)ws, 256));
  GridNmsWorkspace w;
  w.stats = reinterpret_cast<ComboStats*>(p); p += round_up64(b * 9 * (int64_t)sizeof(ComboStats), 256);
  w.meta = reinterpret_cast<uint32_t*>(p); p += round_up64(b * n * 4, 256);
  w.dcount = reinterpret_cast<int32_t*>(p); p += round_up64(b * 4, 256);
  w.dlist = reinterpret_cast<uint16_t*>(p); p += round_up64(b * n * 2, 256);
  w.cnt = reinterpret_cast<uint16_t*>(p); p += round_up64(b * n * 2, 256);
  w.sup = reinterpret_cast<uint16_t*>(p); p += round_up64(b * n * kSupCap * 2, 256);
  w.keep = reinterpret_cast<uint8_t*>(p);
  return w;
}

// anchor index -> (level, cell x, cell y, anchor)
__device__ __forceinline__ void grid_locate(const GridGeom& g, int id, int& l, int& gx, int& gy, int& a) {
  l = id >= g.off[2] ? 2 : (id >= g.off[1] ? 1 : 0);
  const int local = id - g.off[l];
  const int cell = local / 3;
  a = local - cell * 3;
  gy = cell / g.G[l];
  gx = cell - gy * g.G[l];
}

constexpr float kGridEps = 1e-5f;     // slack on every geometric comparison (coordinates are in [0, 1])

// One axis of the regularity test.  lo/hi: box extent, cell [cl, cr].  An unclipped box must have
// its midpoint in its cell; a box clipped at 0 may have its cell anywhere left of the midpoint,
// a box clipped at 1 anywhere right of it (the decoded centre is inside the clipped box, on the
// clipped side of the midpoint).
__device__ __forceinline__ bool grid_axis_regular(float lo, float hi, float cl, float cr) {
  const float mid = 0.5f * (lo + hi);
  const bool need_a = hi < 1.0f;   // not clipped at 1: cell starts at or left of the midpoint
  const bool need_b = lo > 0.0f;   // not clipped at 0: cell ends at or right of the midpoint
  if (need_a && !(cl <= mid + kGridEps)) return false;
  if (need_b && !(cr >= mid - kGridEps)) return false;
  return true;
}

__device__ __forceinline__ bool grid_box_regular(float4 b, int gx, int gy, int G) {
  const float w = b.z - b.x, h = b.w - b.y;
  if (!(w > 1e-12f && h > 1e-12f)) return false;                                  // NaN fails too
  if (!(b.x >= 0.f && b.y >= 0.f && b.z <= 1.f && b.w <= 1.f)) return false;
  const float inv = 1.0f / (float)G;
  return grid_axis_regular(b.x, b.z, gx * inv, (gx + 1) * inv) && grid_axis_regular(b.y, b.w, gy * inv, (gy + 1) * inv);
}

// ---- prep ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
nms_grid_prep_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ cls,
                     GridGeom g, float conf, GridNmsWorkspace w) {
  __shared__ ComboStats s_st[9];
  __shared__ int s_dcount;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 9) {
    s_st[tid].wmin = s_st[tid].hmin = 0x7f800000u;
    s_st[tid].wmax = s_st[tid].hmax = 0u;
    for (int k = 0; k < 8; ++k) s_st[tid].cls[k] = 0u;
  }
  if (tid == 0) s_dcount = 0;
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (int64_t)b * g.N;
  const float* sc = scores + (int64_t)b * g.N;
  const int32_t* cl = cls + (int64_t)b * g.N;
  uint32_t* meta = w.meta + (int64_t)b * g.N;
  uint16_t* dlist = w.dlist + (int64_t)b * g.N;
  for (int id = tid; id < g.N; id += 1024) {
    uint32_t m = 0u;
    if (sc[id] >= conf) {                                   // NaN fails, as in NumPy (:258)
      const uint32_t c = (uint32_t)cl[id] & 0xffu;
      const float4 v = bx[id];
      int l, gx, gy, a;
      grid_locate(g, id, l, gx, gy, a);
      m = c + 1u;
      if (grid_box_regular(v, gx, gy, g.G[l])) {
        ComboStats& st = s_st[l * 3 + a];
        const uint32_t wb = __float_as_uint(v.z - v.x), hb = __float_as_uint(v.w - v.y);
        atomicMin(&st.wmin, wb); atomicMax(&st.wmax, wb);
        atomicMin(&st.hmin, hb); atomicMax(&st.hmax, hb);
        atomicOr(&st.cls[c >> 5], 1u << (c & 31u));
      } else {
        m |= kMetaIrregular;
        dlist[atomicAdd(&s_dcount, 1)] = (uint16_t)id;
      }
    }
    meta[id] = m;
  }
  __syncthreads();
  if (tid < 9) w.stats[(int64_t)b * 9 + tid] = s_st[tid];
  if (tid == 0) w.dcount[b] = s_dcount;
}

// ---- the suppressor search ------------------------------------------------------------------------
struct GridImg {             // one image's arrays
  const float4* boxes;
  const float* scores;
  const uint32_t* meta;
  const ComboStats* stats;   // [9]
  const uint16_t* dlist;
  int dcount;
};

struct GridCtx {             // kernel parameter (constant bank: dynamic indexing is free there)
  GridGeom g;
  IouOps op;
  float t;                   // effective IoU lower bound of a suppressing pair
  float r;                   // |midpoint distance| < r * (own size)
  float ext;                 // a box closer than ext * size to a border may meet partners clipped there
  bool full;                 // threshold too small for any geometric bound: all pairs
};

inline void grid_ctx_threshold(GridCtx& c, float thr, bool diou) {
  c.op = IouOps{thr, diou};
  // suppresses() => computed ovr > thr (or NaN, irregular boxes only) => exact IoU > t
  float t = thr * (1.0f - 1e-5f) - 1e-6f;
  c.full = !(t >= 0.02f);
  if (c.full) t = 0.02f;
  if (t > 1.0f) t = 1.0f;
  c.t = t;
  c.r = fmaxf(1.0f - t, (1.0f - t) / (2.0f * t));
  c.ext = 1.0f / t - t;
}

template <class F>
__device__ __forceinline__ bool grid_visit(const GridCtx& c, const GridImg& im, int i, int j, uint32_t cj, float sj, float4 bj, float aj,
                                           bool regular_only, F& f) {
  const uint32_t mi = im.meta[i];
  if ((mi & kMetaClsMask) != cj) return false;                 // other class / not a candidate
  if (regular_only && (mi & kMetaIrregular)) return false;     // those come through dlist
  const float si = im.scores[i];
  if (!(si > sj || (si == sj && i < j))) return false;         // i must precede j
  const float4 bi = im.boxes[i];
  if (!suppresses(bi, box_area(bi), bj, aj, c.op)) return false;
  return f(i);
}

// Calls f(i) for every candidate i that precedes j, has j's class and suppresses it; stops when
// f returns true.
template <class F>
__device__ void grid_for_each_suppressor(const GridCtx& c, const GridImg& im, int j, uint32_t mj, F& f) {
  const uint32_t cj = mj & kMetaClsMask;
  const float sj = im.scores[j];
  const float4 bj = im.boxes[j];
  const float aj = box_area(bj);
  if ((mj & kMetaIrregular) || c.full) {
    for (int i = 0; i < c.g.N; ++i)
      if (grid_visit(c, im, i, j, cj, sj, bj, aj, false, f)) return;
    return;
  }
  for (int d = 0; d < im.dcount; ++d)
    if (grid_visit(c, im, (int)im.dlist[d], j, cj, sj, bj, aj, false, f)) return;

  const float wj = bj.z - bj.x, hj = bj.w - bj.y;
  const float mx = 0.5f * (bj.x + bj.z), my = 0.5f * (bj.y + bj.w);
  const float e2 = 2.0f * kGridEps;
  float left = mx - c.r * wj - e2, right = mx + c.r * wj + e2;
  float top = my - c.r * hj - e2, bottom = my + c.r * hj + e2;
  if (bj.x < wj * c.ext + e2) left = -e2;             // partners clipped at x = 0: centre anywhere left
  if (1.0f - bj.z < wj * c.ext + e2) right = 1.0f + e2;
  if (bj.y < hj * c.ext + e2) top = -e2;
  if (1.0f - bj.w < hj * c.ext + e2) bottom = 1.0f + e2;
  // partner sizes: t < w_i / w_j < 1 / t
  const float wlo = wj * c.t * (1.0f - 1e-4f), whi = wj / c.t * (1.0f + 1e-4f);
  const float hlo = hj * c.t * (1.0f - 1e-4f), hhi = hj / c.t * (1.0f + 1e-4f);
  const uint32_t cbit = cj - 1u;
#pragma unroll 1
  for (int l = 0; l < 3; ++l) {
    uint32_t amask = 0u;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const ComboStats& st = im.stats[l * 3 + a];
      const bool has_cls = (st.cls[cbit >> 5] >> (cbit & 31u)) & 1u;
      const bool size_ok = __uint_as_float(st.wmax) > wlo && __uint_as_float(st.wmin) < whi &&
                           __uint_as_float(st.hmax) > hlo && __uint_as_float(st.hmin) < hhi;
      if (has_cls && size_ok) amask |= 1u << a;
    }
    if (!amask) continue;
    const int G = c.g.G[l];
    const float Gf = (float)G;
    // cells [k/G, (k+1)/G] that can hold a centre in [left, right]
    const int gx0 = max(0, (int)ceilf(left * Gf - 1.001f)), gx1 = min(G - 1, (int)floorf(right * Gf + 0.001f));
    const int gy0 = max(0, (int)ceilf(top * Gf - 1.001f)), gy1 = min(G - 1, (int)floorf(bottom * Gf + 0.001f));
    for (int gy = gy0; gy <= gy1; ++gy) {
      int i = c.g.off[l] + (gy * G + gx0) * 3;
      for (int gx = gx0; gx <= gx1; ++gx, i += 3) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
          if ((amask >> a) & 1u)
            if (grid_visit(c, im, i + a, j, cj, sj, bj, aj, true, f)) return;
      }
    }
  }
}

__device__ __forceinline__ GridImg grid_image(const GridCtx& proto, const float* boxes, const float* scores,
                                              const GridNmsWorkspace& w, int b) {
  GridImg c;
  const int64_t N = proto.g.N;
  c.boxes = reinterpret_cast<const float4*>(boxes) + (int64_t)b * N;
  c.scores = scores + (int64_t)b * N;
  c.meta = w.meta + (int64_t)b * N;
  c.stats = w.stats + (int64_t)b * 9;
  c.dlist = w.dlist + (int64_t)b * N;
  c.dcount = w.dcount[b];
  return c;
}

// ---- build ----------------------------------------------------------------------------------------
struct SupAppend {
  uint16_t* list;
  int n;
  __device__ __forceinline__ bool operator()(int i) {
    if (n < kSupCap) list[n] = (uint16_t)i;
    ++n;
    return false;
  }
};

// Thread t of an image handles the t-th anchor in (level, anchor, cell) order, so that a warp
// holds 32 neighbouring cells of one anchor shape: similar windows, shared cache lines.
__global__ void __launch_bounds__(256)
nms_grid_build_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                      const __grid_constant__ GridCtx proto, GridNmsWorkspace w) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (t >= proto.g.N) return;
  const GridGeom& g = proto.g;
  const int l = t >= g.off[2] ? 2 : (t >= g.off[1] ? 1 : 0);
  const int local = t - g.off[l], cells = g.G[l] * g.G[l];
  const int a = local / cells, cell = local - a * cells;
  const int j = g.off[l] + cell * 3 + a;
  const int64_t N = g.N;
  const uint32_t mj = w.meta[(int64_t)b * N + j];
  int n = 0;
  if (mj) {
    const GridImg im = grid_image(proto, boxes, scores, w, b);
    SupAppend f{w.sup + ((int64_t)b * N + j) * kSupCap, 0};
    grid_for_each_suppressor(proto, im, j, mj, f);
    n = min(f.n, 65535);
  }
  w.cnt[(int64_t)b * N + j] = (uint16_t)n;
}

// ---- resolve --------------------------------------------------------------------------------------
struct SupRescan {
  const uint8_t* state;
  bool removed, pending;
  __device__ __forceinline__ bool operator()(int i) {
    const uint8_t s = state[i];
    if (s == 2) { removed = true; return true; }
    if (s == 1) pending = true;
    return false;
  }
};

constexpr int kResolveThreads = 1024;

__global__ void __launch_bounds__(kResolveThreads)
nms_grid_resolve_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                        const __grid_constant__ GridCtx proto, GridNmsWorkspace w, uint8_t* __restrict__ keep) {
  extern __shared__ uint8_t s_state[];     // [N]  0 none, 1 undecided, 2 kept, 3 removed
  __shared__ int s_progress;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = proto.g.N;
  pdl_trigger();
  pdl_wait();
  const uint32_t* meta = w.meta + (int64_t)b * N;
  const uint16_t* cnt = w.cnt + (int64_t)b * N;
  const uint16_t* sup = w.sup + (int64_t)b * N * kSupCap;
  for (int j = tid; j < N; j += kResolveThreads) s_state[j] = meta[j] ? (cnt[j] ? 1 : 2) : 0;
  const GridImg im = grid_image(proto, boxes, scores, w, b);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_progress = 0;
    __syncthreads();
    bool progress = false;
    for (int j = tid; j < N; j += kResolveThreads) {
      if (s_state[j] != 1) continue;
      const int n = cnt[j];
      bool removed = false, pending = false;
      if (n <= kSupCap) {
        const uint16_t* lst = sup + (int64_t)j * kSupCap;
        for (int q = 0; q < n; ++q) {
          const uint8_t s = s_state[lst[q]];
          removed |= s == 2;
          pending |= s == 1;
        }
      } else {
        const uint16_t* lst = sup + (int64_t)j * kSupCap;
        for (int q = 0; q < kSupCap; ++q) removed |= s_state[lst[q]] == 2;
        if (!removed) {            // the list is incomplete: evaluate the definition itself
          SupRescan f{s_state, false, false};
          grid_for_each_suppressor(proto, im, j, meta[j], f);
          removed = f.removed;
          pending = f.pending;
        }
      }
      if (removed) { s_state[j] = 3; progress = true; }
      else if (!pending) { s_state[j] = 2; progress = true; }
    }
    if (progress) s_progress = 1;
    __syncthreads();
    if (!s_progress) break;
  }
  uint8_t* kp = keep + (int64_t)b * N;
  for (int j = tid; j < N; j += kResolveThreads) kp[j] = s_state[j] == 2;
}

// ---- host side ---------------------------------------------------------------------------------------
inline cudaError_t launch_nms_grid(const float* boxes, const float* scores, const int32_t* cls, int batch,
                                   int input_size, int num_classes, float conf, float thr, int diou,
                                   float* out_boxes, float* out_scores, int32_t* out_cls, int32_t* out_counts,
                                   uint8_t* keep_out, GridNmsWorkspace w, cudaStream_t st) {
  GridCtx proto{};
  proto.g = make_grid_geom(input_size);
  const int N = proto.g.N;
  if (N > 65535 || num_classes > 255 || input_size % 32) return cudaErrorInvalidValue;
  grid_ctx_threshold(proto, thr, diou != 0);
  uint8_t* keep = keep_out ? keep_out : w.keep;
  cudaError_t r = launch_pdl(nms_grid_prep_kernel, dim3(batch), dim3(1024), 0, st, boxes, scores, cls, proto.g, conf, w);
  if (r != cudaSuccess) return r;
  YNB_COUNT_LAUNCH();
  r = launch_pdl(nms_grid_build_kernel, dim3((N + 255) / 256, batch), dim3(256), 0, st, boxes, scores, proto, w);
  if (r != cudaSuccess) return r;
  YNB_COUNT_LAUNCH();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(nms_grid_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  r = launch_pdl(nms_grid_resolve_kernel, dim3(batch), dim3(kResolveThreads), (size_t)round_up(N, 16), st, boxes, scores,
                 proto, w, keep);
  if (r != cudaSuccess) return r;
  YNB_COUNT_LAUNCH();
  r = launch_pdl(nms_compact_kernel, dim3(batch), dim3(1024), 0, st, (const uint8_t*)keep, boxes, scores, cls, out_boxes,
                 out_scores, out_cls, out_counts, (int64_t)N);
  YNB_COUNT_LAUNCH();
  return r;
}

}  // namespace ynb
