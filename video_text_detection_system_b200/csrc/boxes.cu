// Stage 4a: box extraction on the GPU.
//
// Replaces TextDetector._post_process (text_detector.py:143-178), i.e. OpenCV's
// findContours(RETR_EXTERNAL) -> contourArea -> minAreaRect -> boxPoints plus the reference's own
// truncate / clip / scale / size-filter / mean-probability arithmetic, without leaving the device.
//
// Pipeline per plane (all kernels take a batch of planes in blockIdx.z / a flat frame index):
//   1. ccl_rows      : every pixel gets the raster index of the start of its row run (runs of
//                      foreground AND of background: background is labelled too, 4-connected, so the
//                      RETR_EXTERNAL nesting rule can be decided without tracing hole borders).
//   2. ccl_merge     : one union per pair of touching runs in adjacent rows (8-connectivity for
//                      foreground, 4 for background), lock-free union-find with atomicMin => the root
//                      of a component is its raster-first pixel, which is where cv::findContours
//                      starts the outer border.
//   3. ccl_flatten   : label[i] = root(i).
//   4. mark_outside  : background roots that reach the frame.
//   5. collect_roots : one slot per foreground component (bbox seed, external flag).
//   6. run_extents   : per row run, atomics into the slot's bbox.
//   7. select        : external && bbox can hold area >= 100  -> candidate list.
//   8. geometry      : one thread per candidate: outer border trace (area + per-row extremes), convex
//                      hull, rotating calipers, boxPoints, truncation, AABB, clip, scale, size filter
//                      (csrc/box_geom.cuh, shared with the CPU unit-test harness).
//   9. confidence    : one warp per surviving box, mean of the probability plane over the box.
//  10. pack          : order by raster start index, keep the first kmax, write vtd_record.
// HBM traffic is a handful of passes over the u8 mask / int32 label plane; everything after step 6
// touches O(#components) data.
#include "common.cuh"
#include "box_geom.cuh"
#include "../../include/vtd.h"

namespace vtd {
namespace {

using namespace geom;

__device__ __forceinline__ int uf_find(const int* __restrict__ L, int a) {
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) { int old = atomicMin(L + b, a); done = (old == b); b = old; }
    else if (b < a) { int old = atomicMin(L + a, b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// ---- 1. row runs.  One CTA per (row, plane); thread t owns a contiguous chunk of the row.
constexpr int RT = 256;
__global__ void __launch_bounds__(RT) ccl_rows_kernel(const uint8_t* __restrict__ mask, int* __restrict__ labels,
                                                      int mh, int mw) {
  __shared__ int carry[RT];
  const int y = blockIdx.x;
  const size_t plane = (size_t)blockIdx.y * mh * mw;
  const uint8_t* m = mask + plane + (size_t)y * mw;
  int* L = labels + plane + (size_t)y * mw;
  const int per = (mw + RT - 1) / RT;
  const int x0 = threadIdx.x * per, x1 = min(x0 + per, mw);
  // last run start inside my chunk (or -1 if my chunk continues the run that enters it)
  int last = -1;
  for (int x = x0; x < x1; ++x) {
    bool start = (x == 0) || ((m[x] != 0) != (m[x - 1] != 0));
    if (start) last = x;
  }
  carry[threadIdx.x] = last;
  __syncthreads();
  // inclusive max-scan of `last` over threads
  for (int off = 1; off < RT; off <<= 1) {
    int v = carry[threadIdx.x];
    int o = threadIdx.x >= off ? carry[threadIdx.x - off] : -1;
    __syncthreads();
    carry[threadIdx.x] = max(v, o);
    __syncthreads();
  }
  int cur = threadIdx.x > 0 ? carry[threadIdx.x - 1] : 0;
  for (int x = x0; x < x1; ++x) {
    bool start = (x == 0) || ((m[x] != 0) != (m[x - 1] != 0));
    if (start) cur = x;
    L[x] = y * mw + cur;
  }
}

// ---- 2. merge runs of adjacent rows
// Four pixels per thread: the two mask rows come in as 32-bit words (plus the byte left and right of them), so a pixel
// costs ~1.5 loads instead of 5; the union rules are the per-pixel ones, unchanged.
__device__ __forceinline__ void ccl_merge_px(const uint8_t* __restrict__ m, int* L, int mw, int x, int i, bool f, bool n, bool w,
                                             bool nw, bool ne, bool e) {
  if (f) {
    if (n) {
      if (!w || !nw) uf_union(L, i, i - mw);
    } else {
      if (x > 0 && nw && !w) uf_union(L, i, i - mw - 1);
      if (x + 1 < mw && ne && !e) uf_union(L, i, i - mw + 1);
    }
  } else {
    if (!n) {
      if (w || nw || x == 0) uf_union(L, i, i - mw);
    }
  }
}

__global__ void ccl_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ labels, int mh, int mw) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y + 1;
  if (x0 >= mw || y >= mh) return;
  const size_t plane = (size_t)blockIdx.z * mh * mw;
  const uint8_t* m = mask + plane;
  int* L = labels + plane;
  const int i0 = y * mw + x0;
  uint8_t cur[6], up[6];                                  // pixels x0-1 .. x0+4 of rows y and y-1 (out of frame: see below)
  if ((mw & 3) == 0 && x0 + 4 <= mw) {
    const uint32_t c4 = *reinterpret_cast<const uint32_t*>(m + i0), u4 = *reinterpret_cast<const uint32_t*>(m + i0 - mw);
#pragma unroll
    for (int k = 0; k < 4; ++k) { cur[k + 1] = (uint8_t)(c4 >> (8 * k)); up[k + 1] = (uint8_t)(u4 >> (8 * k)); }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool in = x0 + k < mw;
      cur[k + 1] = in ? m[i0 + k] : 0; up[k + 1] = in ? m[i0 + k - mw] : 0;
    }
  }
  cur[0] = x0 > 0 ? m[i0 - 1] : 0; up[0] = x0 > 0 ? m[i0 - mw - 1] : 0;
  cur[5] = x0 + 4 < mw ? m[i0 + 4] : 0; up[5] = x0 + 4 < mw ? m[i0 + 4 - mw] : 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    if (x >= mw) break;
    const bool f = cur[k + 1] != 0;
    const bool w = x > 0 ? (cur[k] != 0) : !f;              // out of frame == "other class"
    const bool nw = x > 0 ? (up[k] != 0) : !f;
    ccl_merge_px(m, L, mw, x, i0 + k, f, up[k + 1] != 0, w, nw, up[k + 2] != 0, cur[k + 2] != 0);
  }
}
// NOTE on the foreground NE rule: when N is background and NE is foreground, pixel E (if foreground)
// sees NE as its N with a background NW and performs the union itself; only when E is background does
// this pixel have to do it.

// Four pixels per thread (16-byte load/store); pixels of one run carry the same label after ccl_rows/ccl_merge, so a
// run costs one root chase per 4 pixels instead of four.
__global__ void ccl_flatten_plane_kernel(int* __restrict__ labels, int plane_px) {
  int* L = labels + (size_t)blockIdx.y * plane_px;
  if (plane_px & 3) {                                   // planes not 16-byte aligned: scalar
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane_px; i += gridDim.x * blockDim.x) L[i] = uf_find(L, i);
    return;
  }
  const int quads = plane_px >> 2;
  for (int qd = blockIdx.x * blockDim.x + threadIdx.x; qd < quads; qd += gridDim.x * blockDim.x) {
    int4 l = reinterpret_cast<const int4*>(L)[qd];
    int4 r;
    r.x = uf_find(L, l.x);
    r.y = l.y == l.x ? r.x : uf_find(L, l.y);
    r.z = l.z == l.y ? r.y : uf_find(L, l.z);
    r.w = l.w == l.z ? r.z : uf_find(L, l.w);
    reinterpret_cast<int4*>(L)[qd] = r;
  }
}

// ---- 4. background roots connected to the frame
__global__ void mark_outside_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ labels,
                                    uint8_t* __restrict__ outside, int mh, int mw) {
  const size_t plane = (size_t)blockIdx.y * mh * mw;
  const int per = 2 * (mh + mw);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per; t += gridDim.x * blockDim.x) {
    int x, y;
    if (t < mw) { x = t; y = 0; }
    else if (t < 2 * mw) { x = t - mw; y = mh - 1; }
    else if (t < 2 * mw + mh) { x = 0; y = t - 2 * mw; }
    else { x = mw - 1; y = t - 2 * mw - mh; }
    int i = y * mw + x;
    if (mask[plane + i] == 0) outside[plane + labels[plane + i]] = 1;
  }
}

// ---- 5. component slots
struct CompArrays {
  int* start;     // [n][cap] raster index of the component's first pixel
  int* xmin;      // [n][cap]
  int* xmax;
  int* ymax;
  int* external;  // [n][cap]
  int* count;     // [n]
  int cap;
};

__global__ void collect_roots_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ labels,
                                     const uint8_t* __restrict__ outside, int* __restrict__ slot_plane,
                                     CompArrays ca, int mh, int mw) {
  const int f = blockIdx.y;
  const int plane_px = mh * mw;
  const size_t plane = (size_t)f * plane_px;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane_px; i += gridDim.x * blockDim.x) {
    if (mask[plane + i] == 0 || labels[plane + i] != i) continue;
    int s = atomicAdd(ca.count + f, 1);
    slot_plane[plane + i] = s;
    if (s >= ca.cap) continue;
    int x = i % mw, y = i / mw;
    size_t o = (size_t)f * ca.cap + s;
    ca.start[o] = i;
    ca.xmin[o] = x; ca.xmax[o] = x; ca.ymax[o] = y;
    // RETR_EXTERNAL: the pixel left of the raster-first pixel is background; the component is top-level
    // iff that background region reaches the (zero-padded) frame.
    ca.external[o] = (x == 0 || y == 0) ? 1 : (int)outside[plane + labels[plane + i - 1]];
  }
}

// ---- 6. bbox of every component from its row runs
__global__ void run_extents_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ labels,
                                   const int* __restrict__ slot_plane, CompArrays ca, int mh, int mw) {
  const int f = blockIdx.y;
  const int plane_px = mh * mw;
  const size_t plane = (size_t)f * plane_px;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane_px; i += gridDim.x * blockDim.x) {
    if (mask[plane + i] == 0) continue;
    int x = i % mw;
    bool rs = (x == 0) || mask[plane + i - 1] == 0;
    bool re = (x == mw - 1) || mask[plane + i + 1] == 0;
    if (!rs && !re) continue;
    int s = slot_plane[plane + labels[plane + i]];
    if (s >= ca.cap) continue;
    size_t o = (size_t)f * ca.cap + s;
    if (rs && x < ca.xmin[o]) atomicMin(ca.xmin + o, x);
    if (re) {
      if (x > ca.xmax[o]) atomicMax(ca.xmax + o, x);
      int y = i / mw;
      if (y > ca.ymax[o]) atomicMax(ca.ymax + o, y);
    }
  }
}

// ---- 7. candidates
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

// ---- 8. geometry of one candidate
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

// One WARP per candidate.  The candidate's bounding box of the mask (+1 px border) is staged in shared memory, so
// the inherently sequential outer-border trace runs against ~25-cycle shared-memory probes instead of L2 round
// trips; per-row extremes come from the label plane in parallel (lanes stride over x); lane 0 then runs the hull /
// rotating-calipers arithmetic (O(#rows), scratch in shared memory), and the whole warp reduces the mean
// probability of the resulting box.  Components too large for the per-warp budget fall back to global scratch.
#ifdef VTD_TIMERS
__device__ int getenv_timers = 1;
#endif
constexpr int GW = 4;                    // candidates (warps) per CTA
constexpr int GSMEM = 16 * 1024;         // shared-memory bytes per candidate

__global__ void __launch_bounds__(GW * 32) geometry_kernel(const uint8_t* __restrict__ mask,
                                                           const int* __restrict__ labels,
                                                           const float* __restrict__ prob, CompArrays ca, CandArrays cd,
                                                           GeoParams gp, int* __restrict__ pool,
                                                           int* __restrict__ pool_used, TmpBox* __restrict__ tmp,
                                                           int* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y;
  const int nc = min(cd.count[f], cd.kc);
  // a fixed, small grid walks the candidate list (the capacity is 1024 per plane, a frame has ~50: one CTA per
  // capacity slot spent most of the kernel scheduling 4096 CTAs of 64 KB that exit at once)
  for (int c = blockIdx.x * GW + warp; c < nc; c += gridDim.x * GW) {
#ifdef VTD_TIMERS
  const long long gt0 = clock64(); long long gt1 = 0, gt2 = 0, gt3 = 0, gt4 = 0, gt5 = 0;
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
  const int rw = xmax - xmin + 3, rh = nrows + 2;          // staged region incl. 1 px border
  const int bx0 = xmin - 1, by0 = y0 - 1;
  const int mask_bytes = (rw * rh + 15) & ~15;
  const int scratch_words = 12 * nrows + 16;
  const bool use_smem = mask_bytes + 4 * scratch_words <= GSMEM;
  uint8_t* sm_mask = gsm + warp * GSMEM;
  int* scr;
  if (use_smem) {
    scr = reinterpret_cast<int*>(sm_mask + mask_bytes);
  } else {
    int off = 0;
    if (lane == 0) off = atomicAdd(pool_used + f, scratch_words);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + scratch_words > gp.pool_words) { if (lane == 0) atomicExch(overflow, 1); continue; }
    scr = pool + (size_t)f * gp.pool_words + off;
  }
  int* rowmin = scr;
  int* rowmax = scr + nrows;
  Pt* hull = reinterpret_cast<Pt*>(scr + 2 * nrows);
  float* fl = reinterpret_cast<float*>(scr + 2 * nrows + 2 * (2 * nrows + 2));
  const uint8_t* m = mask + (size_t)f * mh * mw;
  const int* L = labels + (size_t)f * mh * mw;

  if (use_smem) {
    // membership of THIS component (label == raster index of its first pixel), bbox + 1 px border.  Foreground pixels
    // 8-adjacent to a pixel of the component belong to the component, so the outer-border trace sees the same
    // neighbourhood in this plane as in the mask.  Rows are fetched 4 at a time: 4 x ceil(rw/32) independent loads in
    // flight per lane instead of one dependent L2 round trip per element.
    for (int yy = 0; yy < rh; yy += 4) {
      for (int xx = lane; xx < rw; xx += 32) {
        const int gx = bx0 + xx;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gy = by0 + yy + k;
          v[k] = (yy + k < rh && (unsigned)gx < (unsigned)mw && (unsigned)gy < (unsigned)mh) ? L[(size_t)gy * mw + gx] : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (yy + k < rh) sm_mask[(yy + k) * rw + xx] = v[k] == start ? (uint8_t)1 : (uint8_t)0;
      }
    }
    __syncwarp();
    // per-row extremes: one lane per row, scanning shared memory
    for (int r = lane; r < nrows; r += 32) {
      const uint8_t* row = sm_mask + (r + 1) * rw + 1;        // x = xmin at offset 0
      const int wd = xmax - xmin + 1;
      int lo = 0, hi = wd - 1;
      while (lo < wd && !row[lo]) ++lo;
      while (hi >= 0 && !row[hi]) --hi;
      rowmin[r] = lo < wd ? xmin + lo : (1 << 30);
      rowmax[r] = hi >= 0 ? xmin + hi : -1;
    }
  } else {
  // per-row extremes of THIS component (label == raster index of its first pixel)
  for (int r = 0; r < nrows; ++r) {
    int lo = 1 << 30, hi = -1;
    const int* row = L + (size_t)(y0 + r) * mw;
    for (int x = xmin + lane; x <= xmax; x += 32)
      if (row[x] == start) { lo = min(lo, x); hi = max(hi, x); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if (lane == 0) { rowmin[r] = lo; rowmax[r] = hi; }
  }
  }
  __syncwarp();

#ifdef VTD_TIMERS
  gt1 = clock64();
#endif
  int ok = 0;
  if (lane == 0) {
    do {
      auto fg_s = [&](int x, int y) -> bool { return sm_mask[(y - by0) * rw + (x - bx0)] != 0; };
      auto fg_g = [&](int x, int y) -> bool {
        return (unsigned)x < (unsigned)mw && (unsigned)y < (unsigned)mh && m[y * mw + x] != 0;
      };
      const long long max_steps = 8LL * mw * mh;
      long long area2 = use_smem ? trace_outer_area2(fg_s, x0, y0, max_steps, nullptr)
                                 : trace_outer_area2(fg_g, x0, y0, max_steps, nullptr);
#ifdef VTD_TIMERS
      gt2 = clock64();
#endif
      if (area2 < 0) area2 = -area2;
      if (area2 < 200) break;                  // cv2.contourArea(contour) < 100 -> skip (text_detector.py:150)
      int nh = hull_from_rows(rowmin, rowmax, y0, nrows, hull);
#ifdef VTD_TIMERS
      gt3 = clock64();
#endif
      if (nh < 3) break;
      RotRect rr = min_area_rect(hull, nh, fl, fl + nh, fl + 2 * nh);
#ifdef VTD_TIMERS
      gt4 = clock64();
#endif
      unclip_rect(rr, gp.unclip);
      PtF bp[4];
      box_points(rr, bp);
      int xs[4], ys[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {             // np.int0: truncation toward zero (text_detector.py:155)
        xs[k] = (int)bp[k].x; ys[k] = (int)bp[k].y;
        tb.poly[2 * k] = xs[k]; tb.poly[2 * k + 1] = ys[k];
      }
      int bx1 = min(min(xs[0], xs[1]), min(xs[2], xs[3])), bx2 = max(max(xs[0], xs[1]), max(xs[2], xs[3]));
      int by1 = min(min(ys[0], ys[1]), min(ys[2], ys[3])), by2 = max(max(ys[0], ys[1]), max(ys[2], ys[3]));
      bx1 = max(0, bx1); by1 = max(0, by1);                       // :160
      bx2 = min(gp.clip_w, bx2); by2 = min(gp.clip_h, by2);       // :161
      int X1 = (int)((double)((long long)bx1 * gp.orig_w) / (double)gp.clip_w);   // :163-166
      int Y1 = (int)((double)((long long)by1 * gp.orig_h) / (double)gp.clip_h);
      int X2 = (int)((double)((long long)bx2 * gp.orig_w) / (double)gp.clip_w);
      int Y2 = (int)((double)((long long)by2 * gp.orig_h) / (double)gp.clip_h);
      if (!(X2 - X1 > 10 && Y2 - Y1 > 10)) break;                 // :168
      tb.bbox[0] = X1; tb.bbox[1] = Y1; tb.bbox[2] = X2; tb.bbox[3] = Y2;
      // :169-170  prob_map[y1*640//oh : y2*640//oh, x1*640//ow : x2*640//ow]  (numpy slice clamps to the plane)
      long long cy0 = (long long)Y1 * gp.clip_h / gp.orig_h, cy1 = (long long)Y2 * gp.clip_h / gp.orig_h;
      long long cx0 = (long long)X1 * gp.clip_w / gp.orig_w, cx1 = (long long)X2 * gp.clip_w / gp.orig_w;
      tb.cy0 = (int)min(cy0, (long long)mh); tb.cy1 = (int)min(cy1, (long long)mh);
      tb.cx0 = (int)min(cx0, (long long)mw); tb.cx1 = (int)min(cx1, (long long)mw);
      tb.start = start;
      ok = 1;
    } while (false);
  }
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
  gt5 = clock64();
  if (lane == 0 && f == 0 && c < 2 && getenv_timers)
    printf("GEOM c=%d rw=%d rh=%d nrows=%d | stage+extremes %lld trace %lld hull %lld calipers %lld rest+conf %lld total %lld\n", c, rw, rh,
           nrows, gt1 - gt0, gt2 - gt1, gt3 - gt2, gt4 - gt3, gt5 - gt4, gt5 - gt0);
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
  const size_t px = (size_t)mh * mw;
  const int cap = ((mh + 1) / 2) * ((mw + 1) / 2) + 1;     // 8-connected components cannot be denser
  const int pool_words = 256 * mh * 12 + 4096;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  lay->cap = cap; lay->kc = kc; lay->pool_words = pool_words;
  lay->labels = take(px * n * 4);
  lay->slot_plane = take(px * n * 4);
  lay->outside = take(px * n);
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
  const size_t px = (size_t)mh * mw;
  int* labels = reinterpret_cast<int*>(work + lay.labels);
  int* slot_plane = reinterpret_cast<int*>(work + lay.slot_plane);
  uint8_t* outside = work + lay.outside;
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
  if ((e = cudaMemsetAsync(outside, 0, px * n, s)) != cudaSuccess) return e;

  ccl_rows_kernel<<<dim3(mh, n), RT, 0, s>>>(mask, labels, mh, mw);
  if (mh > 1) ccl_merge_kernel<<<dim3(cdiv(cdiv(mw, 4), 128), mh - 1, n), 128, 0, s>>>(mask, labels, mh, mw);
  const int gx = min(cdiv((long long)px, 256), 148 * 8);
  ccl_flatten_plane_kernel<<<dim3(min(cdiv((long long)px / 4 + 1, 256), 148 * 8), n), 256, 0, s>>>(labels, (int)px);
  mark_outside_kernel<<<dim3(cdiv(2 * (mh + mw), 256), n), 256, 0, s>>>(mask, labels, outside, mh, mw);
  collect_roots_kernel<<<dim3(gx, n), 256, 0, s>>>(mask, labels, outside, slot_plane, ca, mh, mw);
  run_extents_kernel<<<dim3(gx, n), 256, 0, s>>>(mask, labels, slot_plane, ca, mh, mw);
  select_kernel<<<dim3(min(cdiv(lay.cap, 256), 148), n), 256, 0, s>>>(ca, cd, mw);
  GeoParams gp{mh, mw, p.clip_h, p.clip_w, p.orig_h, p.orig_w, p.unclip, lay.pool_words};
  static PerDeviceFlag geo_attr;
  {
    cudaError_t ge = once_per_device(geo_attr, [] {
      return cudaFuncSetAttribute(geometry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GW * GSMEM);
    });
    if (ge != cudaSuccess) return ge;
  }
  geometry_kernel<<<dim3(min(cdiv(lay.kc, GW), 32), n), GW * 32, GW * GSMEM, s>>>(mask, labels, prob, ca, cd, gp, pool, pool_used,
                                                                        tmp, overflow);
  pack_kernel<<<n, 256, 0, s>>>(cd, tmp, reinterpret_cast<vtd_record*>(records), counts, p.kmax, overflow);
  if (lc) lc->n += (mh > 1 ? 9 : 8);
  return cudaGetLastError();
}

}  // namespace vtd
