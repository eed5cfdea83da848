// Anchor-grid NMS: the product path's replacement for the sort + greedy loop of
// models/yolo_nano.py:159-279, exact (same keep set), without a sort and without a serial scan.
//
// Greedy NMS keeps candidate j iff no KEPT candidate i of the same class that precedes j
// (score descending, ties by ascending anchor index) has ovr(i, j) > thresh (or NaN).  That
// recurrence has exactly one solution, so it can be evaluated in any order:
//
//   prep     one CTA per image: per-anchor meta word (class, candidate, irregular), per
//            (level, anchor) size ranges and class-presence masks, list of irregular candidates
//   build    one thread per candidate; result: every candidate's SUPPRESSOR LIST = the preceding
//            same-class candidates with ovr > thresh (each pair is tested once, from its lower
//            anchor index, and appended to the list of whichever comes later in the order).  Candidates are anchors of a known grid: the decoded centre of
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
//            ... and then compacts the kept candidates in anchor order, the reference's output
//            order (:274-277)
//
// Pair tests are `suppresses()` of decode_nms.cuh, the same IEEE sequence as the reference.
#pragma once
#include "decode_nms.cuh"

namespace ynb {

constexpr int kSupCap = 64;            // listed suppressors per candidate (longer lists: re-scan path)
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

// One anchor, as the search reads it: two 16-byte loads, no dependent chain.
struct alignas(32) GridRec {
  float4 box;
  float area;          // box_area(box), the reference's (x2-x1)*(y2-y1)
  float score;
  uint32_t meta;       // 0 = not a candidate; (class + 1) | kMetaIrregular
  uint32_t pad;
};

struct GridNmsWorkspace {
  ComboStats* stats;   // [B][9]
  GridRec* rec;        // [B][N]
  uint32_t* meta;      // [B][N] (compact copy of rec.meta)
  int32_t* dcount;     // [B]
  uint16_t* dlist;     // [B][N] irregular candidates
  int32_t* cnt;        // [B][N] suppressors found (atomic); > kSupCap: list incomplete
  uint16_t* sup;       // [B][N][kSupCap]
  uint8_t* keep;       // [B][N]
};

inline int64_t grid_nms_workspace_bytes(int batch, int64_t n) {
  int64_t b = batch;
  return round_up64(b * 9 * (int64_t)sizeof(ComboStats), 256) + round_up64(b * n * 32, 256) +
         round_up64(b * n * 4, 256) + round_up64(b * 4, 256) +
         round_up64(b * n * 2, 256) + round_up64(b * n * 4, 256) + round_up64(b * n * kSupCap * 2, 256) +
         round_up64(b * n, 256) + 512;
}
inline GridNmsWorkspace grid_nms_carve(void* ws, int batch, int64_t n) {
  int64_t b = batch;
  char* p = reinterpret_cast<char*>(round_up64((int64_t)(uintptr_t)ws, 256));
  GridNmsWorkspace w;
  w.stats = reinterpret_cast<ComboStats*>(p); p += round_up64(b * 9 * (int64_t)sizeof(ComboStats), 256);
  w.rec = reinterpret_cast<GridRec*>(p); p += round_up64(b * n * 32, 256);
  w.meta = reinterpret_cast<uint32_t*>(p); p += round_up64(b * n * 4, 256);
  w.dcount = reinterpret_cast<int32_t*>(p); p += round_up64(b * 4, 256);
  w.dlist = reinterpret_cast<uint16_t*>(p); p += round_up64(b * n * 2, 256);
  w.cnt = reinterpret_cast<int32_t*>(p); p += round_up64(b * n * 4, 256);
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
  float4* rec = reinterpret_cast<float4*>(w.rec + (int64_t)b * g.N);
  uint16_t* dlist = w.dlist + (int64_t)b * g.N;
  for (int id = tid; id < g.N; id += 1024) {
    uint32_t m = 0u;
    const float score = sc[id];
    const float4 v = bx[id];
    if (score >= conf) {                                    // NaN fails, as in NumPy (:258)
      const uint32_t c = (uint32_t)cl[id] & 0xffu;
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
    w.cnt[(int64_t)b * g.N + id] = 0;
    rec[2 * id] = v;
    rec[2 * id + 1] = make_float4(box_area(v), score, __uint_as_float(m), 0.0f);
  }
  __syncthreads();
  if (tid < 9) w.stats[(int64_t)b * 9 + tid] = s_st[tid];
  if (tid == 0) w.dcount[b] = s_dcount;
}

// ---- the suppressor search ------------------------------------------------------------------------
struct GridImg {             // one image's arrays
  const float4* rec;         // GridRec as two float4
  const ComboStats* stats;   // [9] (shared-memory copy in the build kernel)
  const uint16_t* dlist;
  int dcount;
};

struct GridCtx {             // kernel parameter (constant bank: dynamic indexing is free there)
  GridGeom g;
  IouOps op;
  float t;                   // effective IoU lower bound of a suppressing pair
  float r;                   // |midpoint distance| < r * (own size)
  float ext;                 // only a box closer than ext * size to a border can meet a partner clipped
                             // there: [0, w_i] and [x1, x1 + w] share > t max(w_i, w)  =>  x1 < w (1 - t) / t
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
  c.ext = (1.0f - t) / t;
}

struct GridSelf {            // the candidate whose suppressors are searched
  int j;
  int level, gx, gy;         // own cell (half-window mode only)
  uint32_t cls1;             // class + 1
  float score, area;
  float4 box;
};

// Both halves of the record are loaded unconditionally (independent loads, neighbouring lanes
// share the lines); class, order and the IoU test are evaluated without early exits so the
// warp stays converged — only a hit (rare) branches into f(i, i_precedes_j).
//   kHalf = false: hits are the candidates that PRECEDE j and suppress it (pull mode)
//   kHalf = true : hits are the candidates with a LARGER anchor index that form a suppressing pair
//                  with j in either order — every regular pair is examined once, from its lower index
template <bool kHalf, class F>
__device__ __forceinline__ bool grid_visit(const GridCtx& c, const GridImg& im, int i, const GridSelf& me,
                                           bool regular_only, F& f) {
  const float4 bi = im.rec[2 * i];
  const float4 qi = im.rec[2 * i + 1];
  const uint32_t mi = __float_as_uint(qi.z);
  const bool prec = qi.y > me.score || (qi.y == me.score && i < me.j);     // i precedes j
  bool pass = (mi & kMetaClsMask) == me.cls1;                              // same class, candidate
  pass &= !(regular_only && (mi & kMetaIrregular));                        // those come through dlist
  pass &= kHalf ? i > me.j : prec;
  pass &= suppresses(bi, qi.x, me.box, me.area, c.op);
  return pass ? f(i, prec) : false;
}

// Calls f for every hit of candidate j (see grid_visit); stops when f returns true.  In half-window
// mode irregular candidates / irregular partners / the all-pairs fallback still work in pull mode
// (each side collects its own predecessors), only regular-regular pairs are split by index.
template <bool kHalf, class F>
__device__ void grid_for_each_suppressor(const GridCtx& c, const GridImg& im, GridSelf& me, F& f) {
  const int j = me.j;
  me.box = im.rec[2 * j];
  const float4 qj = im.rec[2 * j + 1];
  me.area = qj.x;
  me.score = qj.y;
  const uint32_t mj = __float_as_uint(qj.z);
  me.cls1 = mj & kMetaClsMask;
  if ((mj & kMetaIrregular) || c.full) {
    for (int i = 0; i < c.g.N; ++i)
      if (grid_visit<false>(c, im, i, me, false, f)) return;
    return;
  }
  for (int d = 0; d < im.dcount; ++d)
    if (grid_visit<false>(c, im, (int)im.dlist[d], me, false, f)) return;

  const float4 bj = me.box;
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
  const uint32_t cbit = me.cls1 - 1u;
  // (level, anchor) combinations that can hold a partner — a converged loop, no window work inside
  uint32_t kmask = 0u;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const ComboStats& st = im.stats[k];
    const bool has_cls = (st.cls[cbit >> 5] >> (cbit & 31u)) & 1u;
    const bool size_ok = __uint_as_float(st.wmax) > wlo && __uint_as_float(st.wmin) < whi &&
                         __uint_as_float(st.hmax) > hlo && __uint_as_float(st.hmin) < hhi;
    if (has_cls && size_ok) kmask |= 1u << k;
  }
  if (kHalf) kmask &= ~((1u << (me.level * 3)) - 1u);          // lower levels hold lower indices only
  // One loop over all windows: every lane walks its own (combination, row, column) state, so the
  // warp stays converged on the visit no matter how the lanes' windows differ.
  int i = 0, gx = 0, gy = 0, gx0 = 0, gx1 = -1, gy1 = -1, G = 1, rowi = 0;
  bool have = false;
  for (;;) {
    if (!have) {
      if (!kmask) break;
      const int k = __ffs((int)kmask) - 1;
      kmask &= kmask - 1u;
      const int l = k / 3, a = k - l * 3;
      G = c.g.G[l];
      const float Gf = (float)G;
      // cells [q/G, (q+1)/G] that can hold a centre in [left, right]
      gx0 = max(0, (int)ceilf(left * Gf - 1.001f));
      gx1 = min(G - 1, (int)floorf(right * Gf + 0.001f));
      int gy0 = max(0, (int)ceilf(top * Gf - 1.001f));
      gy1 = min(G - 1, (int)floorf(bottom * Gf + 0.001f));
      const bool own_level = kHalf && l == me.level;
      if (own_level) gy0 = max(gy0, me.gy);                      // earlier rows hold lower indices
      if (gx0 > gx1 || gy0 > gy1) continue;
      gy = gy0;
      gx = own_level && gy == me.gy ? max(gx0, me.gx) : gx0;    // own row: earlier columns likewise
      if (gx > gx1) {
        if (++gy > gy1) continue;
        gx = gx0;
      }
      rowi = c.g.off[l] + a;
      i = rowi + (gy * G + gx) * 3;
      have = true;
    }
    if (grid_visit<kHalf>(c, im, i, me, true, f)) return;
    ++gx;
    i += 3;
    if (gx > gx1) {
      ++gy;
      gx = gx0;
      i = rowi + (gy * G + gx0) * 3;
      have = gy <= gy1;
    }
  }
}

__device__ __forceinline__ GridImg grid_image(const GridCtx& proto, const GridNmsWorkspace& w, int b) {
  GridImg c;
  const int64_t N = proto.g.N;
  c.rec = reinterpret_cast<const float4*>(w.rec + (int64_t)b * N);
  c.stats = w.stats + (int64_t)b * 9;
  c.dlist = w.dlist + (int64_t)b * N;
  c.dcount = w.dcount[b];
  return c;
}

// ---- build ----------------------------------------------------------------------------------------
struct SupAppend {          // hit (i, j): the later one of the pair gets the earlier one in its list
  int32_t* cnt;             // this image
  uint16_t* sup;
  int j;
  __device__ __forceinline__ bool operator()(int i, bool i_first) {
    const int later = i_first ? j : i, earlier = i_first ? i : j;
    const int pos = atomicAdd(cnt + later, 1);
    if (pos < kSupCap) sup[(int64_t)later * kSupCap + pos] = (uint16_t)earlier;
    return false;
  }
};

// Thread t of an image handles the t-th anchor in (level, anchor, cell) order, so that a warp
// holds 32 neighbouring cells of one anchor shape: similar windows, shared cache lines.
__global__ void __launch_bounds__(256)
nms_grid_build_kernel(const __grid_constant__ GridCtx proto, GridNmsWorkspace w) {
  __shared__ ComboStats s_stats[9];
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x < 9 * (int)(sizeof(ComboStats) / 4))
    reinterpret_cast<uint32_t*>(s_stats)[threadIdx.x] =
        reinterpret_cast<const uint32_t*>(w.stats + (int64_t)b * 9)[threadIdx.x];
  __syncthreads();
  if (t >= proto.g.N) return;
  const GridGeom& g = proto.g;
  const int l = t >= g.off[2] ? 2 : (t >= g.off[1] ? 1 : 0);
  const int local = t - g.off[l], cells = g.G[l] * g.G[l];
  const int a = local / cells, cell = local - a * cells;
  const int64_t N = g.N;
  GridSelf me;
  me.j = g.off[l] + cell * 3 + a;
  me.level = l;
  me.gy = cell / g.G[l];
  me.gx = cell - me.gy * g.G[l];
  if (!w.meta[(int64_t)b * N + me.j]) return;
  GridImg im = grid_image(proto, w, b);
  im.stats = s_stats;
  SupAppend f{w.cnt + (int64_t)b * N, w.sup + (int64_t)b * N * kSupCap, me.j};
  grid_for_each_suppressor<true>(proto, im, me, f);
}

// ---- resolve --------------------------------------------------------------------------------------
struct SupRescan {
  const uint8_t* state;
  bool removed, pending;
  __device__ __forceinline__ bool operator()(int i, bool) {
    const uint8_t s = state[i];
    if (s == 2) { removed = true; return true; }
    if (s == 1) pending = true;
    return false;
  }
};

constexpr int kResolveThreads = 1024;

// One CTA per image.  Shared memory: state[N] (0 none, 1 undecided, 2 kept, 3 removed), cnt8[N]
// (suppressor count, 255 = "more than the list holds") and — when list_cap > 0 — the suppressor
// lists themselves with their offsets, so that a sweep costs shared-memory latency only.  Thread t
// owns candidates t, t+1024, ...; a 64-bit register mask tracks which of them are still
// undecided, so a sweep only touches those.  After the fixed point the same CTA compacts the
// kept candidates in anchor order (models/yolo_nano.py:274-277).
constexpr int kResolveSmemBytes = 200 * 1024;

__global__ void __launch_bounds__(kResolveThreads)
nms_grid_resolve_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                        const int32_t* __restrict__ cls, const __grid_constant__ GridCtx proto, GridNmsWorkspace w,
                        uint8_t* __restrict__ keep, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                        int32_t* __restrict__ out_cls, int32_t* __restrict__ out_counts, int list_cap) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ int s_progress;
  __shared__ int s_warp[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = proto.g.N, Np = (N + 15) & ~15;
  uint8_t* s_state = s_dyn;
  uint8_t* s_cnt = s_dyn + Np;
  uint32_t* s_off = reinterpret_cast<uint32_t*>(s_dyn + 2 * Np);         // [Np], only if list_cap > 0
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_dyn + 6 * Np);        // [list_cap]
  pdl_trigger();
  pdl_wait();
  const uint32_t* meta = w.meta + (int64_t)b * N;
  const int32_t* cnt = w.cnt + (int64_t)b * N;
  const uint16_t* sup = w.sup + (int64_t)b * N * kSupCap;
  int mine = 0;                                    // list entries of this thread's undecided candidates
  unsigned long long und = 0ull;                   // bit k: candidate tid + k*1024 is undecided
  {
    int k = 0;
    for (int j = tid; j < N; j += kResolveThreads, ++k) {
      const int n = cnt[j];
      const bool cand = meta[j] != 0u;
      s_cnt[j] = (uint8_t)(n > kSupCap ? 255 : n);
      s_state[j] = cand ? (n ? 1 : 2) : 0;
      if (cand && n) {
        mine += min(n, kSupCap);
        und |= 1ull << k;
      }
    }
  }
  bool in_smem = false;
  if (list_cap > 0) {
    // exclusive scan of `mine` over the CTA -> this thread's run [base, base + mine)
    int incl0 = mine;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl0, d);
      if ((tid & 31) >= d) incl0 += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl0;
    __syncthreads();
    if (tid < 32) {
      int wv = s_warp[tid], wi = wv;
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, wi, d);
        if (tid >= d) wi += v;
      }
      s_warp[tid] = wi - wv;
    }
    __syncthreads();
    const int base = s_warp[tid >> 5] + incl0 - mine;
    in_smem = base + mine <= list_cap;
    if (in_smem) {
      // copy this thread's lists, four candidates at a time so that their first 16-byte parts
      // (8 entries: most lists) are in flight together
      int off = base;
      auto put8 = [&](uint4 q, int dst, int n) {
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int h = 0; h < 8; ++h)
          if (h < n) s_list[dst + h] = (uint16_t)(wd[h >> 1] >> ((h & 1) * 16));
      };
      unsigned long long m = und;
      while (m) {
        int jj[4], nn[4];
        uint4 q0[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          jj[u] = -1;
          nn[u] = 0;
          if (m) {
            jj[u] = tid + (__ffsll((long long)m) - 1) * kResolveThreads;
            m &= m - 1;
            nn[u] = min((int)s_cnt[jj[u]], kSupCap);
            q0[u] = *reinterpret_cast<const uint4*>(sup + (int64_t)jj[u] * kSupCap);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (jj[u] < 0) continue;
          s_off[jj[u]] = (uint32_t)off;
          put8(q0[u], off, nn[u]);
          const uint4* src = reinterpret_cast<const uint4*>(sup + (int64_t)jj[u] * kSupCap);
          for (int v = 1; v * 8 < nn[u]; ++v) put8(src[v], off + v * 8, nn[u] - v * 8);
          off += nn[u];
        }
      }
    }
  }
  const GridImg im = grid_image(proto, w, b);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_progress = 0;
    __syncthreads();
    bool progress = false;
    for (unsigned long long m = und; m; m &= m - 1) {
      const int k = __ffsll((long long)m) - 1;
      const int j = tid + k * kResolveThreads;
      const int n = s_cnt[j];
      const int listed = min(n, kSupCap);
      bool removed = false, pending = false;
      if (in_smem) {
        const uint16_t* lst = s_list + s_off[j];
        for (int q = 0; q < listed; ++q) {
          const uint8_t s = s_state[lst[q]];
          removed |= s == 2;
          pending |= s == 1;
        }
      } else {
        const uint16_t* lst = sup + (int64_t)j * kSupCap;
        for (int q = 0; q < listed; ++q) {
          const uint8_t s = s_state[lst[q]];
          removed |= s == 2;
          pending |= s == 1;
        }
      }
      if (n == 255 && !removed && !pending) {   // list incomplete and exhausted: evaluate the definition itself
        SupRescan f{s_state, false, false};
        GridSelf me;
        me.j = j;
        me.level = me.gx = me.gy = 0;
        grid_for_each_suppressor<false>(proto, im, me, f);
        removed = f.removed;
        pending = f.pending;
      }
      if (removed || !pending) {
        s_state[j] = removed ? 3 : 2;
        und &= ~(1ull << k);
        progress = true;
      }
    }
    if (progress) s_progress = 1;
    __syncthreads();
    if (!s_progress) break;
  }
  // ---- compaction, anchor order: warp w owns anchors [w * per, (w + 1) * per), 32 at a time
  const int lane = tid & 31, wid = tid >> 5;
  const int per = ((N + 31) / 32 + 31) & ~31;                 // multiple of 32 anchors per warp
  const int beg = min(N, wid * per), end = min(N, beg + per);
  uint8_t* kp = keep + (int64_t)b * N;
  int wtotal = 0;
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool k = i < end && s_state[i] == 2;
    if (i < end) kp[i] = k;
    wtotal += __popc(__ballot_sync(0xffffffffu, k));
  }
  __syncthreads();                                            // s_warp is free again
  if (lane == 0) s_warp[wid] = wtotal;
  __syncthreads();
  if (tid < 32) {
    int wv = s_warp[tid], wi = wv;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (tid >= d) wi += v;
    }
    s_warp[tid] = wi - wv;
    if (tid == 31) out_counts[b] = wi;
  }
  __syncthreads();
  int pos0 = s_warp[wid];
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (int64_t)b * N;
  const float* sc = scores + (int64_t)b * N;
  const int32_t* cl = cls + (int64_t)b * N;
  float4* ob = reinterpret_cast<float4*>(out_boxes) + (int64_t)b * N;
  float* os = out_scores + (int64_t)b * N;
  int32_t* oc = out_cls + (int64_t)b * N;
#pragma unroll 4
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool k = i < end && s_state[i] == 2;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int pos = pos0 + __popc(bal & ((1u << lane) - 1u));
      ob[pos] = bx[i];
      os[pos] = sc[i];
      oc[pos] = cl[i];
    }
    pos0 += __popc(bal);
  }
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
  r = launch_pdl(nms_grid_build_kernel, dim3((N + 255) / 256, batch), dim3(256), 0, st, proto, w);
  if (r != cudaSuccess) return r;
  YNB_COUNT_LAUNCH();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(nms_grid_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kResolveSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  // shared-memory lists need 6 bytes per anchor for state / count / offset; beyond that the lists stay in L2
  const int Np = round_up(N, 16);
  const int list_cap = 6 * Np + 32768 <= kResolveSmemBytes ? (kResolveSmemBytes - 6 * Np) / 2 : 0;
  r = launch_pdl(nms_grid_resolve_kernel, dim3(batch), dim3(kResolveThreads), (size_t)kResolveSmemBytes, st, boxes, scores,
                 cls, proto, w, keep, out_boxes, out_scores, out_cls, out_counts, list_cap);
  YNB_COUNT_LAUNCH();
  return r;
}

}  // namespace ynb
