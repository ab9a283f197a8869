// Per-component geometry of the box extraction stage, written once for host and device.
//
// The reference's _post_process (text_detector.py:143-178) leans on four OpenCV calls whose source is
// not part of the reference tree (opencv-python-headless, requirements.txt:15): findContours
// (RETR_EXTERNAL, Suzuki-Abe border following), contourArea (shoelace over the border polygon),
// minAreaRect (convex hull + rotating calipers in float32) and boxPoints.  The functions below restate
// those published algorithms so that the results agree with OpenCV on integer pixel sets.  They are
// compiled into the CUDA kernels (boxes.cu) and, for the CPU unit tests only, into a host harness
// (tests/host_harness.cpp) that is compared against cv2 itself.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VTD_HD __host__ __device__ __forceinline__
#else
#define VTD_HD inline
#endif

namespace vtd {
namespace geom {

// float32 arithmetic exactly as written (no FMA contraction), on both sides
#if defined(__CUDA_ARCH__)
VTD_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
VTD_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
VTD_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
VTD_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
VTD_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
#else
VTD_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
VTD_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
VTD_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
VTD_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
VTD_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
#endif

struct Pt { int x, y; };
struct PtF { float x, y; };

// ---- outer border following ---------------------------------------------------------------------
// Traces the outer border of the 8-connected component whose raster-first pixel is (x0,y0), in the
// order cv::findContours visits it, and returns twice the signed shoelace area of the polygon
// through the visited pixel centres (cv::contourArea = |area2|/2).  Fg(x,y) must return false
// outside the image.  steps_out (optional) receives the number of border steps.
// The 8 directions (E, NE, N, NW, W, SW, S, SE) as packed 2-bit fields (value + 1): no table in local memory on the
// device, where this loop is one lane's dependent chain.
VTD_HD int trace_dx(int s) { return (int)((0x901Au >> (2 * (s & 7))) & 3u) - 1; }   // {1,1,0,-1,-1,-1,0,1}
VTD_HD int trace_dy(int s) { return (int)((0xA901u >> (2 * (s & 7))) & 3u) - 1; }   // {0,-1,-1,-1,0,1,1,1}

// Bit helpers (device intrinsics / portable host loops).
VTD_HD int lowest_set(unsigned v) {              // index of the lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  int i = 0; while (!((v >> i) & 1u)) ++i; return i;
#endif
}
VTD_HD int highest_set(unsigned v) {             // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  int i = 31; while (!((v >> i) & 1u)) --i; return i;
#endif
}

// The trace works on NEIGHBOURHOOD CODES: nbr(x, y) returns 8 bits, bit s set iff the neighbour of (x, y) in direction s
// is foreground (pixels outside the image are background).  One code per border pixel replaces the probe-by-probe search
// for the next border pixel -- on the device the code comes from three bit-rows of a shared-memory window (csrc/boxes.cu),
// six independent loads instead of a chain of dependent ones.
// `visit(x, y)` is called once for every border pixel in visiting order (a pixel the border passes twice is visited
// twice): the box extraction kernel collects the per-row extremes of the component from it (every row extreme of an
// 8-connected component lies on its outer border).
template <class Nbr, class Visit>
VTD_HD long long trace_outer_nbr(const Nbr& nbr, int x0, int y0, long long max_steps, long long* steps_out, const Visit& visit) {
  if (steps_out) *steps_out = 0;
  unsigned code = nbr(x0, y0);
  // first neighbour in the order NW, N, NE, E, SE, S, SW (directions 3, 2, 1, 0, 7, 6, 5); W is never consulted:
  // rotate the code so that direction 3 is bit 7 ... direction 4 is bit 0, and take the highest set bit above bit 0
  const unsigned r0 = (((code << 4) | (code >> 4)) & 0xFEu);
  if (r0 == 0) { visit(x0, y0); return 0; }      // isolated pixel
  int s = (highest_set(r0) - 4) & 7;
  const int x1 = x0 + trace_dx(s), y1 = y0 + trace_dy(s);
  long long area2 = 0, steps = 0;
  int x3 = x0, y3 = y0;
  for (;;) {
    visit(x3, y3);
    // next border pixel: the first foreground neighbour in the order s+1, s+2, ... (the pixel we came from is one)
    const unsigned rot = ((code | (code << 8)) >> ((s + 1) & 7)) & 0xFFu;
    s = (s + 1 + lowest_set(rot)) & 7;
    const int ddx = trace_dx(s), ddy = trace_dy(s);
    const int x4 = x3 + ddx, y4 = y3 + ddy;
    area2 += (long long)(x3 * ddy - ddx * y3);          // == x3*y4 - x4*y3 (consecutive border pixels are 8-neighbours)
    ++steps;
    if ((x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) || steps >= max_steps) break;
    x3 = x4; y3 = y4;
    s = (s + 4) & 7;
    code = nbr(x3, y3);
  }
  if (steps_out) *steps_out = steps;
  return area2;
}

// neighbourhood code from a pixel predicate (host harness; device fallback on the global mask)
template <class Fg>
struct NbrFromFg {
  const Fg& fg;
  VTD_HD unsigned operator()(int x, int y) const {
    unsigned n = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < 8; ++s) n |= (fg(x + trace_dx(s), y + trace_dy(s)) ? 1u : 0u) << s;
    return n;
  }
};

template <class Fg, class Visit>
VTD_HD long long trace_outer_visit(const Fg& fg, int x0, int y0, long long max_steps, long long* steps_out, const Visit& visit) {
  const NbrFromFg<Fg> nbr{fg};
  return trace_outer_nbr(nbr, x0, y0, max_steps, steps_out, visit);
}

struct NoVisit { VTD_HD void operator()(int, int) const {} };

template <class Fg>
VTD_HD long long trace_outer_area2(const Fg& fg, int x0, int y0, long long max_steps, long long* steps_out) {
  return trace_outer_visit(fg, x0, y0, max_steps, steps_out, NoVisit());
}

// ---- convex hull from per-row extremes ------------------------------------------------------------
// rowmin[i], rowmax[i] are the smallest / largest x of the component in row y0+i (i < nrows; every row
// of an 8-connected component is populated).  Every hull vertex is a row extreme, so the hull of the
// 2*nrows extremes is the hull of the component (and of its outer border).  Output: the strictly
// convex vertices in the cyclic order cv::convexHull(points, clockwise=false) returns them: starting
// at the right-most point (largest y among ties), then towards larger y (down the image), the
// left-most point, the top, and back.  `hull` needs room for 2*nrows+2 points.  Returns the count.
// 32-bit on purpose (the chain below is one lane's dependent sequence on the device): exact while the coordinates stay
// below 2^15, which extract_boxes() enforces (planes of at most 16384 x 16384 pixels)
VTD_HD int cross3(const Pt& o, const Pt& a, const Pt& b) {
  return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x);
}

VTD_HD int hull_from_rows(const int* rowmin, const int* rowmax, int y0, int nrows, Pt* hull) {
  // Right chain: rows top -> bottom using row maxima; left chain: bottom -> top using row minima.
  // With x right / y down, walking top->bottom along the right side then bottom->top along the left
  // side is a walk with positive shoelace sign; a vertex is kept when the turn is strictly "right-hand"
  // in that sense (cross > 0).
  int n = 0;
  for (int i = 0; i < nrows; ++i) {
    Pt p = {rowmax[i], y0 + i};
    while (n >= 2 && cross3(hull[n - 2], hull[n - 1], p) <= 0) --n;
    hull[n++] = p;
  }
  int lower = n + 1;
  for (int i = nrows - 1; i >= 0; --i) {
    Pt p = {rowmin[i], y0 + i};
    if (n > 0 && hull[n - 1].x == p.x && hull[n - 1].y == p.y) continue;
    while (n >= lower && cross3(hull[n - 2], hull[n - 1], p) <= 0) --n;
    hull[n++] = p;
  }
  // close: drop the last point if it repeats the first, and fix collinearity across the seam
  if (n > 1 && hull[n - 1].x == hull[0].x && hull[n - 1].y == hull[0].y) --n;
  bool changed = true;
  while (changed && n > 2) {
    changed = false;
    if (cross3(hull[n - 2], hull[n - 1], hull[0]) <= 0) { --n; changed = true; continue; }
    if (cross3(hull[n - 1], hull[0], hull[1]) <= 0) {
      for (int i = 0; i + 1 < n; ++i) hull[i] = hull[i + 1];
      --n; changed = true;
    }
  }
  if (n < 3) return n;
  // cv::minAreaRect calls convexHull(points, clockwise=false): the walk built above (top -> right -> bottom ->
  // left on screen) is already that direction.  OpenCV then shifts the sequence cyclically so that the
  // indices into the contour descend; the contour starts at the component's raster-first pixel, so the
  // vertex (min y, then min x) comes LAST.  (Verified against cv2 4.13 on >6000 blobs; contours that pass
  // twice through 1-pixel-wide appendages can start elsewhere, which only matters for exact ties.)
  int best = 0;
  for (int i = 1; i < n; ++i)
    if (hull[i].y < hull[best].y || (hull[i].y == hull[best].y && hull[i].x < hull[best].x)) best = i;
  const int shift = (best + 1) % n;      // new first element
  if (shift != 0) {
    auto rev = [&](int a, int b) { while (a < b) { Pt t = hull[a]; hull[a] = hull[b]; hull[b] = t; ++a; --b; } };
    rev(0, shift - 1); rev(shift, n - 1); rev(0, n - 1);
  }
  return n;
}

// ---- rotating calipers (cv::minAreaRect on a convex polygon), float32 like OpenCV ---------------------
struct RotRect { float cx, cy, w, h, angle; };

// inv_len / vx / vy are caller scratch of n floats each.
VTD_HD RotRect min_area_rect(const Pt* hp, int n, float* inv_len, float* vx, float* vy) {
  RotRect rr = {0.f, 0.f, 0.f, 0.f, 0.f};
  float minarea = 3.402823466e+38f;
  int left = 0, bottom = 0, right = 0, top = 0;
  float left_x, right_x, top_y, bottom_y;
  float p0x = (float)hp[0].x, p0y = (float)hp[0].y;
  left_x = right_x = p0x; top_y = bottom_y = p0y;
  for (int i = 0; i < n; ++i) {
    if (p0x < left_x) { left_x = p0x; left = i; }
    if (p0x > right_x) { right_x = p0x; right = i; }
    if (p0y > top_y) { top_y = p0y; top = i; }
    if (p0y < bottom_y) { bottom_y = p0y; bottom = i; }
    int j = (i + 1 < n) ? i + 1 : 0;
    float px = (float)hp[j].x, py = (float)hp[j].y;
    double ddx = (double)px - (double)p0x, ddy = (double)py - (double)p0y;
    vx[i] = (float)ddx; vy[i] = (float)ddy;
    inv_len[i] = (float)(1.0 / sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy))));
    p0x = px; p0y = py;
  }
  float orientation = 0.f;
  {
    double ax = vx[n - 1], ay = vy[n - 1];
    for (int i = 0; i < n; ++i) {
      double bx = vx[i], by = vy[i];
      double convexity = dmul(ax, by) - dmul(ay, bx);
      if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
      ax = bx; ay = by;
    }
  }
  float base_a = orientation, base_b = 0.f;
  int seq[4] = {bottom, right, top, left};
  // best-so-far record
  int best_left = 0, best_bottom = 0;
  float best_a = 0.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
  for (int k = 0; k < n; ++k) {
    // OpenCV >= 4.5: pick the caliper whose (rotated) polygon edge is right-most, by cross-product sign
    float rvx[4], rvy[4];
    rvx[0] = vx[seq[0]];  rvy[0] = vy[seq[0]];
    rvx[1] = vy[seq[1]];  rvy[1] = -vx[seq[1]];     // rotate90CW
    rvx[2] = -vx[seq[2]]; rvy[2] = -vy[seq[2]];     // rotate180
    rvx[3] = -vy[seq[3]]; rvy[3] = vx[seq[3]];      // rotate90CCW
    int main_element = 0;
    for (int i = 1; i < 4; ++i) {
      // firstVecIsRight(rv[i], rv[main]): rotate90CW(v1) . v2 < 0
      float tx = rvy[i], ty = -rvx[i];
      if (fadd(fmul(tx, rvx[main_element]), fmul(ty, rvy[main_element])) < 0.f) main_element = i;
    }
    {
      int pindex = seq[main_element];
      float lead_x = fmul(vx[pindex], inv_len[pindex]);
      float lead_y = fmul(vy[pindex], inv_len[pindex]);
      switch (main_element) {
        case 0: base_a = lead_x; base_b = lead_y; break;
        case 1: base_a = lead_y; base_b = -lead_x; break;
        case 2: base_a = -lead_x; base_b = -lead_y; break;
        default: base_a = -lead_y; base_b = lead_x; break;
      }
    }
    seq[main_element] += 1;
    if (seq[main_element] == n) seq[main_element] = 0;
    {
      float dx = fsub((float)hp[seq[1]].x, (float)hp[seq[3]].x);
      float dy = fsub((float)hp[seq[1]].y, (float)hp[seq[3]].y);
      float width = fadd(fmul(dx, base_a), fmul(dy, base_b));
      dx = fsub((float)hp[seq[2]].x, (float)hp[seq[0]].x);
      dy = fsub((float)hp[seq[2]].y, (float)hp[seq[0]].y);
      float height = fadd(fmul(-dx, base_b), fmul(dy, base_a));
      float area = fmul(width, height);
      if (area <= minarea) {
        minarea = area;
        best_left = seq[3]; best_a = base_a; best_w = width; best_b = base_b; best_h = height;
        best_bottom = seq[0];
      }
    }
  }
  float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
  float C1 = fadd(fmul(A1, (float)hp[best_left].x), fmul((float)hp[best_left].y, B1));
  float C2 = fadd(fmul(A2, (float)hp[best_bottom].x), fmul((float)hp[best_bottom].y, B2));
  float idet = 1.f / fsub(fmul(A1, B2), fmul(A2, B1));
  float px = fmul(fsub(fmul(C1, B2), fmul(C2, B1)), idet);
  float py = fmul(fsub(fmul(A1, C2), fmul(A2, C1)), idet);
  float o1x = fmul(A1, best_w), o1y = fmul(B1, best_w);
  float o2x = fmul(A2, best_h), o2y = fmul(B2, best_h);
  rr.cx = fadd(px, fmul(fadd(o1x, o2x), 0.5f));
  rr.cy = fadd(py, fmul(fadd(o1y, o2y), 0.5f));
  rr.w = (float)sqrt(dadd(dmul((double)o1x, (double)o1x), dmul((double)o1y, (double)o1y)));
  rr.h = (float)sqrt(dadd(dmul((double)o2x, (double)o2x), dmul((double)o2y, (double)o2y)));
  // angle in degrees, computed in double and folded into [-90, 0) before the single rounding to float
  // (cv2 4.13 behaviour, verified bit-exact: >= 90 -> -180 keeping the sides, else -90 swapping them)
  double ang = atan2((double)o1y, (double)o1x) * 180.0 / 3.1415926535897932384626433832795;
  if (ang >= 90.0) ang -= 180.0;
  else if (ang >= 0.0 || ang < -90.0) { float t = rr.w; rr.w = rr.h; rr.h = t; ang += (ang >= 0.0 ? -90.0 : 90.0); }
  rr.angle = (float)ang;
  return rr;
}

// cv::boxPoints / RotatedRect::points
VTD_HD void box_points(const RotRect& r, PtF pt[4]) {
  double ang = (double)r.angle * 3.1415926535897932384626433832795 / 180.0;
  float b = fmul((float)cos(ang), 0.5f);
  float a = fmul((float)sin(ang), 0.5f);
  pt[0].x = fsub(fsub(r.cx, fmul(a, r.h)), fmul(b, r.w));
  pt[0].y = fsub(fadd(r.cy, fmul(b, r.h)), fmul(a, r.w));
  pt[1].x = fsub(fadd(r.cx, fmul(a, r.h)), fmul(b, r.w));
  pt[1].y = fsub(fsub(r.cy, fmul(b, r.h)), fmul(a, r.w));
  pt[2].x = fsub(fmul(2.f, r.cx), pt[0].x);
  pt[2].y = fsub(fmul(2.f, r.cy), pt[0].y);
  pt[3].x = fsub(fmul(2.f, r.cx), pt[1].x);
  pt[3].y = fsub(fmul(2.f, r.cy), pt[1].y);
}

// north_star's "unclip" (an extension; the reference has none): grow a rotated rect DB-style by the published offset
// d = area * ratio / perimeter on every side.  ratio <= 1 disables it (reference behaviour).  Explicit float32 operations
// (no FMA contraction), op for op as oracle/port.py unclip_rect states them.
VTD_HD void unclip_rect(RotRect& r, float ratio) {
  if (!(ratio > 1.0f)) return;
  const float per = fmul(2.f, fadd(r.w, r.h));
  if (!(per > 0.f)) return;
  const float d = fmul(fmul(r.w, r.h), ratio) / per;
  const float two_d = fmul(2.f, d);
  r.w = fadd(r.w, two_d); r.h = fadd(r.h, two_d);
}

}  // namespace geom
}  // namespace vtd
