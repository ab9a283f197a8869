// Stage 4a: box extraction on the GPU.
//
// Replaces TextDetector._post_process (text_detector.py:143-178), i.e. OpenCV's
// findContours(RETR_EXTERNAL) -> contourArea -> minAreaRect -> boxPoints plus the reference's own
// truncate / clip / scale / size-filter / mean-probability arithmetic, without leaving the device.
//
// Run-length connected components: a text mask is a few dozen blobs on an empty background, i.e. a few runs per row,
// so nothing here is per-pixel except ONE read of the u8 mask -- there is no label plane.
//   1. runs      : one warp per row turns the mask row into its list of runs (start x of every run of equal pixels,
//                  foreground AND background: background is labelled too, 4-connected, so the RETR_EXTERNAL nesting
//                  rule can be decided without tracing hole borders).  Run id = row * cap_row + index.
//   2. merge     : one warp per row; every run looks up the runs of the row above that touch it (binary search in
//                  that row's sorted starts; 8-connectivity for foreground, 4 for background) and unites with the ones
//                  of its colour: lock-free union-find with atomicMin => the root of a component is its raster-first
//                  run, whose first pixel is where cv::findContours starts the outer border.
//   3. flatten   : par[run] = root; background runs on the frame mark their root "outside"; every foreground root
//                  takes a component slot (bbox seeded with its own run).
//   4. extents   : every foreground run folds its extent into its component's bbox; the root decides "external"
//                  (the background run left of the first pixel belongs to a region that reaches the frame).
//   5. select    : external && bbox can hold area >= 100  -> candidate list.
//   6. geometry  : one warp per candidate: the candidate's window of the mask is bit-packed into shared memory; the
//                  outer border trace (area + per-row extremes), convex hull, rotating calipers, boxPoints, truncation,
//                  AABB, clip, scale and size filter run on one lane (csrc/box_geom.cuh, shared with the CPU unit-test
//                  harness); the warp then reduces the mean probability over the box.
//   7. pack      : order by raster start index, keep the first kmax, write vtd_record.
// HBM traffic: the mask once (1 B/px) plus O(#runs); everything after step 1 touches O(#runs) data.
#include "common.cuh"
#include "box_geom.cuh"
#include "../../include/vtd.h"

namespace vtd {
namespace {

using namespace geom;

constexpr int ID_MASK = 0x3fffffff;      // run id inside a plane (< 2^30)
constexpr int OUT_BIT = 0x40000000;      // on a background ROOT: the region touches the frame
constexpr int RW = 8;                    // rows (warps) per CTA of the row kernels

struct RunTabs {
  uint16_t* run_x;   // [n][mh][cap_row] start x of every run, ascending
  int* nruns;        // [n][mh] (count << 1) | colour of run 0 (1 = foreground); colours alternate along a row
  int* par;          // [n][mh][cap_row] union-find parent (plane-relative run id); roots may carry OUT_BIT
  int* slot_of;      // [n][mh][cap_row] component slot, written at foreground roots only
  int cap_row;       // mw + 1
};

__device__ __forceinline__ int uf_find(const int* par, int a) {
  int p = __ldcg(par + a) & ID_MASK;
  while (p != a) { a = p; p = __ldcg(par + a) & ID_MASK; }
  return a;
}

// roots only ever decrease (atomicMin), so a stale read is still an ancestor and the loop below repairs itself
__device__ __forceinline__ void uf_union(int* par, int a, int b) {
  bool done;
  do {
    a = uf_find(par, a);
    b = uf_find(par, b);
    if (a < b) { int old = atomicMin(par + b, a); done = (old == b); b = old; }
    else if (b < a) { int old = atomicMin(par + a, b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// ---- 1. mask row -> runs.  Lane l of the warp takes the 32-pixel words l, l+32, ... of the row; a word's run starts are
// the set bits of  bits ^ (bits << 1 | last bit of the previous word); a warp scan of the counts gives every start its slot.
__device__ __forceinline__ uint32_t mask_word(const uint8_t* __restrict__ m, int x0, int mw, bool vec) {
  uint32_t bits = 0;
  if (vec && x0 + 32 <= mw) {
    const uint4* q = reinterpret_cast<const uint4*>(m + x0);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t v = w[i];
      v |= v >> 4; v |= v >> 2; v |= v >> 1;                     // any bit of a byte -> its bit 0
      bits |= ((((v & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << (4 * i);
    }
  } else {
    for (int i = 0; i < 32 && x0 + i < mw; ++i) bits |= (m[x0 + i] != 0 ? 1u : 0u) << i;
  }
  return bits;
}

__global__ void __launch_bounds__(RW * 32) runs_kernel(const uint8_t* __restrict__ mask, RunTabs rt, int mh, int mw) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * RW + (threadIdx.x >> 5), f = blockIdx.y;
  if (y >= mh) return;
  const size_t row = (size_t)f * mh + y;
  const uint8_t* m = mask + row * mw;
  uint16_t* rx = rt.run_x + row * rt.cap_row;
  int* par = rt.par + row * rt.cap_row;
  const bool vec = (mw & 15) == 0;
  const int words = (mw + 31) >> 5;
  const int id0 = y * rt.cap_row;
  int base = 0;
  uint32_t carry = 0;                                            // last pixel of the previous word (warp-uniform)
  int c0 = 0;
  for (int w0 = 0; w0 < words; w0 += 32) {
    const int w = w0 + lane;
    const bool live = w < words;
    const uint32_t bits = live ? mask_word(m, w * 32, mw, vec) : 0u;
    const int nvalid = live ? min(32, mw - w * 32) : 0;
    uint32_t prev = __shfl_up_sync(0xffffffffu, bits >> 31, 1);
    if (lane == 0) prev = carry;
    if (w == 0) { prev = (~bits) & 1u; c0 = (int)(bits & 1u); }  // x = 0 always starts a run
    uint32_t t = bits ^ ((bits << 1) | (prev & 1u));
    if (nvalid < 32) t &= nvalid > 0 ? ((1u << nvalid) - 1u) : 0u;
    const int cnt = __popc(t);
    int inc = cnt;                                               // inclusive scan over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    int k = base + inc - cnt;
    while (t) {
      const int b = __ffs(t) - 1;
      t &= t - 1;
      rx[k] = (uint16_t)(w * 32 + b);
      par[k] = id0 + k;
      ++k;
    }
    base += __shfl_sync(0xffffffffu, inc, 31);
    // last VALID pixel of the round's last live word (words beyond the row are dead lanes)
    const int last_lane = min(31, words - 1 - w0);
    const uint32_t lb = (bits >> ((nvalid > 0 ? nvalid : 1) - 1)) & 1u;
    carry = __shfl_sync(0xffffffffu, lb, last_lane);
  }
  c0 = __shfl_sync(0xffffffffu, c0, 0);
  if (lane == 0) rt.nruns[row] = (base << 1) | c0;
}

// ---- 2. unite the runs of row y with the touching runs of row y-1
__global__ void __launch_bounds__(RW * 32) merge_kernel(RunTabs rt, int mh, int mw) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * RW + (threadIdx.x >> 5) + 1, f = blockIdx.y;
  if (y >= mh) return;
  const size_t row = (size_t)f * mh + y;
  const int nc_ = rt.nruns[row], nu_ = rt.nruns[row - 1];
  const int nc = nc_ >> 1, c0 = nc_ & 1, nu = nu_ >> 1, u0 = nu_ & 1;
  const uint16_t* xs = rt.run_x + row * rt.cap_row;
  const uint16_t* xu = xs - rt.cap_row;
  int* par = rt.par + (size_t)f * mh * rt.cap_row;
  for (int k = lane; k < nc; k += 32) {
    const int fg = c0 ^ (k & 1);
    const int s = xs[k], e = (k + 1 < nc ? (int)xs[k + 1] : mw) - 1;
    const int lo = fg ? max(s - 1, 0) : s, hi = fg ? min(e + 1, mw - 1) : e;   // 8-connected foreground, 4-connected background
    int a = 0, b = nu - 1;                                       // the run of the row above that holds pixel lo
    while (a < b) { const int mid = (a + b + 1) >> 1; if ((int)xu[mid] <= lo) a = mid; else b = mid - 1; }
    int j = a;
    if ((u0 ^ (j & 1)) != fg) ++j;                               // colours alternate: every second run is mine
    for (; j < nu && (int)xu[j] <= hi; j += 2) uf_union(par, y * rt.cap_row + k, (y - 1) * rt.cap_row + j);
  }
}

struct CompArrays {
  int* start;     // [n][cap] raster index of the component's first pixel
  int* xmin;      // [n][cap]
  int* xmax;
  int* ymax;
  int* external;  // [n][cap]
  int* count;     // [n]
  int cap;
};

// ---- 3. flatten; frame-touching background roots; component slots
__global__ void __launch_bounds__(RW * 32) flatten_kernel(RunTabs rt, CompArrays ca, int mh, int mw) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * RW + (threadIdx.x >> 5), f = blockIdx.y;
  if (y >= mh) return;
  const size_t row = (size_t)f * mh + y;
  const int nc_ = rt.nruns[row];
  const int nc = nc_ >> 1, c0 = nc_ & 1;
  const uint16_t* xs = rt.run_x + row * rt.cap_row;
  const size_t plane = (size_t)f * mh * rt.cap_row;
  int* par = rt.par + plane;
  for (int k = lane; k < nc; k += 32) {
    const int id = y * rt.cap_row + k;
    const int fg = c0 ^ (k & 1);
    const int s = xs[k], e = (k + 1 < nc ? (int)xs[k + 1] : mw) - 1;
    const int r = uf_find(par, id);
    if (r != id) par[id] = r;                                    // roots keep their word (it may carry OUT_BIT)
    if (!fg) {
      if (y == 0 || y == mh - 1 || s == 0 || e == mw - 1) atomicOr(par + r, OUT_BIT);
    } else if (r == id) {
      const int slot = atomicAdd(ca.count + f, 1);
      rt.slot_of[plane + id] = slot;
      if (slot < ca.cap) {
        const size_t o = (size_t)f * ca.cap + slot;
        ca.start[o] = y * mw + s;
        ca.xmin[o] = s; ca.xmax[o] = e; ca.ymax[o] = y;
      }
    }
  }
}

// ---- 4. bbox of every component from its runs; RETR_EXTERNAL flag
__global__ void __launch_bounds__(RW * 32) extents_kernel(RunTabs rt, CompArrays ca, int mh, int mw) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * RW + (threadIdx.x >> 5), f = blockIdx.y;
  if (y >= mh) return;
  const size_t row = (size_t)f * mh + y;
  const int nc_ = rt.nruns[row];
  const int nc = nc_ >> 1, c0 = nc_ & 1;
  const uint16_t* xs = rt.run_x + row * rt.cap_row;
  const size_t plane = (size_t)f * mh * rt.cap_row;
  const int* par = rt.par + plane;
  for (int kk = lane; 2 * kk + (c0 ^ 1) < nc; kk += 32) {        // foreground runs: k = c0 ? 0,2,4.. : 1,3,5..
    const int k = 2 * kk + (c0 ^ 1);
    const int id = y * rt.cap_row + k;
    const int s = xs[k], e = (k + 1 < nc ? (int)xs[k + 1] : mw) - 1;
    const int r = __ldcg(par + id) & ID_MASK;                    // flattened: the root itself
    const int slot = rt.slot_of[plane + r];
    if (slot >= ca.cap) continue;
    const size_t o = (size_t)f * ca.cap + slot;
    if (r != id) {
      if (s < ca.xmin[o]) atomicMin(ca.xmin + o, s);
      if (e > ca.xmax[o]) atomicMax(ca.xmax + o, e);
      if (y > ca.ymax[o]) atomicMax(ca.ymax + o, y);
    } else {
      // RETR_EXTERNAL: the pixel left of the raster-first pixel is background; the component is top-level iff that
      // background region reaches the (zero-padded) frame.
      int ext = 1;
      if (s > 0 && y > 0) {
        const int pl = __ldcg(par + id - 1);
        const int rl = pl & ID_MASK;
        const int word = rl == id - 1 ? pl : __ldcg(par + rl);
        ext = (word & OUT_BIT) ? 1 : 0;
      }
      ca.external[o] = ext;
    }
  }
}

// ---- 5. candidates
struct CandArrays {
  int* slot;      // [n][kc]
  int* count;     // [n]
  int kc;
};

__global__ void select_kernel(CompArrays ca, CandArrays cd, int mw) {
  const int f = blockIdx.y;
  const int nc = min(ca.count[f], ca.cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nc; s += gridDim.x * blockDim.x) {
    size_t o = (size_t)f * ca.cap + s;
    if (!ca.external[o]) continue;
    int y0 = ca.start[o] / mw;
    long long bw = ca.xmax[o] - ca.xmin[o], bh = ca.ymax[o] - y0;
    if (bw * bh < 100) continue;                  // the contour polygon lies inside its (w-1)x(h-1) box
    int c = atomicAdd(cd.count + f, 1);
    if (c < cd.kc) cd.slot[(size_t)f * cd.kc + c] = s;
  }
}

// ---- 6. geometry of one candidate
struct TmpBox {
  int valid;
  int start;
  int bbox[4];
  int poly[8];
  int cy0, cy1, cx0, cx1;   // confidence window (clamped to the plane); empty => NaN
  float conf;
};

struct GeoParams {
  int mh, mw, clip_h, clip_w, orig_h, orig_w;
  float unclip;
  int pool_words;           // per-plane scratch pool size, 4-byte words
};

// One WARP per candidate.  The candidate's window of the mask (bbox + 1 px, columns rounded out to 32) is bit-packed into
// shared memory, one word per 32 pixels; the inherently sequential outer-border trace then takes ONE neighbourhood code
// per border pixel from three bit-rows (six independent shared-memory loads + a funnel shift each) instead of probing
// neighbour after neighbour; the trace itself yields the per-row extremes; lane 0 then runs the hull / rotating-calipers
// arithmetic (O(#rows), scratch in shared memory), and the whole warp reduces the mean probability of the resulting
// box.  Components too large for the per-warp budget fall back to the global mask and global scratch.
constexpr int GW = 4;                    // candidates (warps) per CTA
constexpr int GSMEM = 16 * 1024;         // shared-memory bytes per candidate

struct NbrBits {                          // neighbourhood code from the bit-packed window (each row ends in a zero word)
  const uint32_t* bits; int bx0, by0, wpr;
  __device__ __forceinline__ unsigned operator()(int x, int y) const {
    const int xx = x - bx0 - 1;           // window bit of pixel x-1
    const int sh = xx & 31;
    const uint32_t* r = bits + (y - by0 - 1) * wpr + (xx >> 5);       // row y-1
    const uint32_t t0 = r[0], t1 = r[1], m0 = r[wpr], m1 = r[wpr + 1], b0 = r[2 * wpr], b1 = r[2 * wpr + 1];
    const uint32_t t = __funnelshift_r(t0, t1, sh) & 7u, m = __funnelshift_r(m0, m1, sh) & 7u, b = __funnelshift_r(b0, b1, sh) & 7u;
    // directions: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE; bit 0 of t/m/b is column x-1, bit 2 is x+1
    return (m >> 2) | ((t >> 2) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) | ((m & 1u) << 4) | (b << 5);
  }
};

// Everything one candidate needs after its window is staged; runs on lane 0.  Returns 1 when the box survives.
template <class Nbr>
__device__ __forceinline__ int candidate_geometry(const Nbr& nbr, int* rowmin, int* rowmax, Pt* hull, float* fl, int x0, int y0,
                                                  int nrows, int start, const GeoParams& gp, TmpBox& tb) {
  const int mw = gp.mw, mh = gp.mh;
  auto visit = [&](int x, int y) {
    const int r = y - y0;
    if (x < rowmin[r]) rowmin[r] = x;
    if (x > rowmax[r]) rowmax[r] = x;
  };
  long long area2 = trace_outer_nbr(nbr, x0, y0, 8LL * mw * mh, nullptr, visit);
  if (area2 < 0) area2 = -area2;
  if (area2 < 200) return 0;                  // cv2.contourArea(contour) < 100 -> skip (text_detector.py:150)
  const int nh = hull_from_rows(rowmin, rowmax, y0, nrows, hull);
  if (nh < 3) return 0;
  RotRect rr = min_area_rect(hull, nh, fl, fl + nh, fl + 2 * nh);
  unclip_rect(rr, gp.unclip);
  PtF bp[4];
  box_points(rr, bp);
  int xs[4], ys[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {               // np.int0: truncation toward zero (text_detector.py:155)
    xs[k] = (int)bp[k].x; ys[k] = (int)bp[k].y;
    tb.poly[2 * k] = xs[k]; tb.poly[2 * k + 1] = ys[k];
  }
  int bx1 = min(min(xs[0], xs[1]), min(xs[2], xs[3])), bx2 = max(max(xs[0], xs[1]), max(xs[2], xs[3]));
  int by1 = min(min(ys[0], ys[1]), min(ys[2], ys[3])), by2 = max(max(ys[0], ys[1]), max(ys[2], ys[3]));
  bx1 = max(0, bx1); by1 = max(0, by1);                       // :160
  bx2 = min(gp.clip_w, bx2); by2 = min(gp.clip_h, by2);       // :161
  const int X1 = (int)((double)((long long)bx1 * gp.orig_w) / (double)gp.clip_w);   // :163-166
  const int Y1 = (int)((double)((long long)by1 * gp.orig_h) / (double)gp.clip_h);
  const int X2 = (int)((double)((long long)bx2 * gp.orig_w) / (double)gp.clip_w);
  const int Y2 = (int)((double)((long long)by2 * gp.orig_h) / (double)gp.clip_h);
  if (!(X2 - X1 > 10 && Y2 - Y1 > 10)) return 0;              // :168
  tb.bbox[0] = X1; tb.bbox[1] = Y1; tb.bbox[2] = X2; tb.bbox[3] = Y2;
  // :169-170  prob_map[y1*640//oh : y2*640//oh, x1*640//ow : x2*640//ow]  (numpy slice clamps to the plane)
  const long long cy0 = (long long)Y1 * gp.clip_h / gp.orig_h, cy1 = (long long)Y2 * gp.clip_h / gp.orig_h;
  const long long cx0 = (long long)X1 * gp.clip_w / gp.orig_w, cx1 = (long long)X2 * gp.clip_w / gp.orig_w;
  tb.cy0 = (int)min(cy0, (long long)mh); tb.cy1 = (int)min(cy1, (long long)mh);
  tb.cx0 = (int)min(cx0, (long long)mw); tb.cx1 = (int)min(cx1, (long long)mw);
  tb.start = start;
  return 1;
}

__global__ void __launch_bounds__(GW * 32) geometry_kernel(const uint8_t* __restrict__ mask,
                                                           const float* __restrict__ prob, CompArrays ca, CandArrays cd,
                                                           GeoParams gp, int* __restrict__ pool,
                                                           int* __restrict__ pool_used, TmpBox* __restrict__ tmp,
                                                           int* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y;
  const int nc = min(cd.count[f], cd.kc);
  // a fixed, small grid walks the candidate list (the capacity is 1024 per plane, a frame has ~50)
  for (int c = blockIdx.x * GW + warp; c < nc; c += gridDim.x * GW) {
#ifdef VTD_TIMERS
  const long long gt0 = clock64(); long long gt1 = 0, gt2 = 0;
#endif
  TmpBox& tb = tmp[(size_t)f * cd.kc + c];
  if (lane == 0) tb.valid = 0;
  const int s = cd.slot[(size_t)f * cd.kc + c];
  const size_t o = (size_t)f * ca.cap + s;
  const int start = ca.start[o];
  const int mw = gp.mw, mh = gp.mh;
  const int x0 = start % mw, y0 = start / mw;
  const int xmin = ca.xmin[o], xmax = ca.xmax[o];
  const int nrows = ca.ymax[o] - y0 + 1;
  const int by0 = y0 - 1, rh = nrows + 2;
  const int bx0 = (xmin - 1) & ~31;                          // window origin, a multiple of 32 (-32 when xmin == 0)
  const int wpr = ((xmax + 1 - bx0) >> 5) + 2;               // words per window row, the last one always zero
  const long long words = (long long)wpr * rh;
  const int scratch_words = 12 * nrows + 16;
  const bool use_smem = words * 4 + 4LL * scratch_words <= GSMEM;
  const uint8_t* m = mask + (size_t)f * mh * mw;
  int ok = 0;
  if (use_smem) {
    uint32_t* bitsm = reinterpret_cast<uint32_t*>(gsm + warp * GSMEM);
    int* scr = reinterpret_cast<int*>(bitsm + words);
    int* rowmin = scr;
    int* rowmax = scr + nrows;
    Pt* hull = reinterpret_cast<Pt*>(scr + 2 * nrows);
    float* fl = reinterpret_cast<float*>(scr + 2 * nrows + 2 * (2 * nrows + 2));
    const bool vec = (mw & 15) == 0;
    for (int idx = lane; idx < (int)words; idx += 32) {
      const int r = idx / wpr, wi = idx - r * wpr;
      const int gy = by0 + r, gx = bx0 + 32 * wi;
      uint32_t wbits = 0;
      if (wi + 1 < wpr && (unsigned)gy < (unsigned)mh && gx >= 0 && gx < mw) wbits = mask_word(m + (size_t)gy * mw, gx, mw, vec);
      bitsm[idx] = wbits;
    }
    for (int r = lane; r < nrows; r += 32) { rowmin[r] = 1 << 30; rowmax[r] = -1; }
    __syncwarp();
#ifdef VTD_TIMERS
    gt1 = clock64();
#endif
    if (lane == 0) ok = candidate_geometry(NbrBits{bitsm, bx0, by0, wpr}, rowmin, rowmax, hull, fl, x0, y0, nrows, start, gp, tb);
  } else {
    int off = 0;
    if (lane == 0) off = atomicAdd(pool_used + f, scratch_words);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + scratch_words > gp.pool_words) { if (lane == 0) atomicExch(overflow, 1); continue; }
    int* scr = pool + (size_t)f * gp.pool_words + off;
    int* rowmin = scr;
    int* rowmax = scr + nrows;
    Pt* hull = reinterpret_cast<Pt*>(scr + 2 * nrows);
    float* fl = reinterpret_cast<float*>(scr + 2 * nrows + 2 * (2 * nrows + 2));
    for (int r = lane; r < nrows; r += 32) { rowmin[r] = 1 << 30; rowmax[r] = -1; }
    __syncwarp();
    auto fg_g = [&](int x, int y) -> bool {
      return (unsigned)x < (unsigned)mw && (unsigned)y < (unsigned)mh && m[(size_t)y * mw + x] != 0;
    };
    if (lane == 0) ok = candidate_geometry(NbrFromFg<decltype(fg_g)>{fg_g}, rowmin, rowmax, hull, fl, x0, y0, nrows, start, gp, tb);
  }
#ifdef VTD_TIMERS
  gt2 = clock64();
#endif
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) continue;
  // mean probability inside the box (np.mean of the slice; empty slice -> nan)
  const int cy0 = __shfl_sync(0xffffffffu, tb.cy0, 0), cy1 = __shfl_sync(0xffffffffu, tb.cy1, 0);
  const int cx0 = __shfl_sync(0xffffffffu, tb.cx0, 0), cx1 = __shfl_sync(0xffffffffu, tb.cx1, 0);
  const int h = cy1 - cy0, w = cx1 - cx0;
  float res;
  if (h <= 0 || w <= 0) {
    res = __int_as_float(0x7fc00000);
  } else {
    const float* p = prob + (size_t)f * mh * mw;
    double acc = 0.0;
    for (int y = cy0; y < cy1; y += 4) {                  // 4 rows of loads in flight; same summation order per row
      float rs[4] = {0.f, 0.f, 0.f, 0.f};
      for (int x = cx0 + lane; x < cx1; x += 32) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = y + k < cy1 ? p[(size_t)(y + k) * mw + x] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) rs[k] += v[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) if (y + k < cy1) acc += (double)rs[k];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    res = (float)(acc / ((double)h * (double)w));
  }
#ifdef VTD_TIMERS
  if (lane == 0 && f == 0 && c < 3)
    printf("GEOM c=%d wpr=%d rh=%d nrows=%d smem=%d | stage %lld lane0 %lld conf %lld total %lld\n", c, wpr, rh, nrows, (int)use_smem,
           gt1 - gt0, gt2 - gt1, clock64() - gt2, clock64() - gt0);
#endif
  if (lane == 0) { tb.conf = res; __threadfence_block(); tb.valid = 1; }
  __syncwarp();
  }
}

// ---- 10. order by raster start, keep kmax, write records.  One CTA per plane.
__global__ void pack_kernel(CandArrays cd, const TmpBox* __restrict__ tmp, vtd_record* __restrict__ records,
                            int* __restrict__ counts, int kmax, int* __restrict__ overflow) {
  const int f = blockIdx.x;
  const int nc = min(cd.count[f], cd.kc);
  if (cd.count[f] > cd.kc && threadIdx.x == 0) atomicExch(overflow, 1);
  const TmpBox* t = tmp + (size_t)f * cd.kc;
  __shared__ int nvalid;
  if (threadIdx.x == 0) nvalid = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    if (!t[c].valid) continue;
    int rank = 0;
    const int st = t[c].start;
    for (int j = 0; j < nc; ++j) rank += (t[j].valid && t[j].start < st) ? 1 : 0;
    atomicAdd(&nvalid, 1);
    if (rank >= kmax) continue;
    vtd_record r;
    r.frame = f;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.bbox[k] = t[c].bbox[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) r.polygon[k] = t[c].poly[k];
    r.det_conf = t[c].conf;
    r.rec_conf = 0.f;
    r.len = 0;
#pragma unroll
    for (int k = 0; k < 36; ++k) r.ids[k] = 0;
    r.start_index = st;
#pragma unroll
    for (int k = 0; k < 24; ++k) r.pad[k] = 0;
    records[(size_t)f * kmax + rank] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (nvalid > kmax) atomicExch(overflow, 2);
    counts[f] = min(nvalid, kmax);
  }
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

size_t box_work_bytes(int n, int mh, int mw, int kc, BoxWorkLayout* lay) {
  const int cap = ((mh + 1) / 2) * ((mw + 1) / 2) + 1;     // 8-connected components cannot be denser
  const int pool_words = 256 * mh * 12 + 4096;
  const int cap_row = mw + 1;                               // a row of mw pixels has at most mw runs
  const size_t nrun = (size_t)n * mh * cap_row;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  lay->cap = cap; lay->kc = kc; lay->pool_words = pool_words; lay->cap_row = cap_row;
  lay->run_x = take(nrun * 2);
  lay->nruns = take((size_t)n * mh * 4);
  lay->par = take(nrun * 4);
  lay->slot_of = take(nrun * 4);
  lay->comp = take((size_t)n * cap * 4 * 5);
  lay->cand_slot = take((size_t)n * kc * 4);
  lay->tmp = take((size_t)n * kc * sizeof(TmpBox));
  lay->pool = take((size_t)n * pool_words * 4);
  lay->zero_begin = off;
  lay->comp_count = take((size_t)n * 4);
  lay->cand_count = take((size_t)n * 4);
  lay->pool_used = take((size_t)n * 4);
  lay->overflow = take(4);
  lay->zero_end = off;
  return off;
}

cudaError_t extract_boxes(const float* prob, const uint8_t* mask, const BoxParams& p, uint8_t* work,
                          const BoxWorkLayout& lay, void* records, int* counts, cudaStream_t s,
                          LaunchCounter* lc) {
  if (p.n <= 0) return cudaSuccess;
  const int n = p.n, mh = p.mh, mw = p.mw;
  if (mw > 16384 || mh > 16384 || (long long)mh * lay.cap_row > ID_MASK) return cudaErrorInvalidValue;   // 32-bit hull arithmetic, 16-bit run starts
  RunTabs rt;
  rt.run_x = reinterpret_cast<uint16_t*>(work + lay.run_x);
  rt.nruns = reinterpret_cast<int*>(work + lay.nruns);
  rt.par = reinterpret_cast<int*>(work + lay.par);
  rt.slot_of = reinterpret_cast<int*>(work + lay.slot_of);
  rt.cap_row = lay.cap_row;
  CompArrays ca;
  int* comp = reinterpret_cast<int*>(work + lay.comp);
  const size_t cs = (size_t)p.n_alloc * lay.cap;
  ca.start = comp; ca.xmin = comp + cs; ca.xmax = comp + 2 * cs; ca.ymax = comp + 3 * cs; ca.external = comp + 4 * cs;
  ca.count = reinterpret_cast<int*>(work + lay.comp_count);
  ca.cap = lay.cap;
  CandArrays cd;
  cd.slot = reinterpret_cast<int*>(work + lay.cand_slot);
  cd.count = reinterpret_cast<int*>(work + lay.cand_count);
  cd.kc = lay.kc;
  TmpBox* tmp = reinterpret_cast<TmpBox*>(work + lay.tmp);
  int* pool = reinterpret_cast<int*>(work + lay.pool);
  int* pool_used = reinterpret_cast<int*>(work + lay.pool_used);
  int* overflow = reinterpret_cast<int*>(work + lay.overflow);
  cudaError_t e;
  if ((e = cudaMemsetAsync(work + lay.zero_begin, 0, lay.zero_end - lay.zero_begin, s)) != cudaSuccess) return e;

  const dim3 rows(cdiv(mh, RW), n);
  runs_kernel<<<rows, RW * 32, 0, s>>>(mask, rt, mh, mw);
  if (mh > 1) merge_kernel<<<dim3(cdiv(mh - 1, RW), n), RW * 32, 0, s>>>(rt, mh, mw);
  flatten_kernel<<<rows, RW * 32, 0, s>>>(rt, ca, mh, mw);
  extents_kernel<<<rows, RW * 32, 0, s>>>(rt, ca, mh, mw);
  select_kernel<<<dim3(min(cdiv(lay.cap, 256), 148), n), 256, 0, s>>>(ca, cd, mw);
  GeoParams gp{mh, mw, p.clip_h, p.clip_w, p.orig_h, p.orig_w, p.unclip, lay.pool_words};
  static PerDeviceFlag geo_attr;
  {
    cudaError_t ge = once_per_device(geo_attr, [] {
      return cudaFuncSetAttribute(geometry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GW * GSMEM);
    });
    if (ge != cudaSuccess) return ge;
  }
  geometry_kernel<<<dim3(min(cdiv(lay.kc, GW), 32), n), GW * 32, GW * GSMEM, s>>>(mask, prob, ca, cd, gp, pool, pool_used, tmp,
                                                                                  overflow);
  pack_kernel<<<n, 256, 0, s>>>(cd, tmp, reinterpret_cast<vtd_record*>(records), counts, p.kmax, overflow);
  if (lc) lc->n += (mh > 1 ? 7 : 6);
  return cudaGetLastError();
}

}  // namespace vtd
