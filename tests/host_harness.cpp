// TEST-ONLY host harness: runs the product's host/device geometry header (csrc/box_geom.cuh) on the
// CPU so the contour / hull / min-area-rect logic can be compared with cv2 without a GPU.  It is built
// by tests/conftest.py into tests/_build/ and is never linked into libvtd_b200.so.
//
// Labelling here is a plain sequential flood fill (the GPU uses union-find, boxes.cu); the rule that
// decides RETR_EXTERNAL membership and everything per component is the shared header.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../video_text_detection_system_b200/csrc/box_geom.cuh"
#include "../video_text_detection_system_b200/csrc/resize_tab.h"

using namespace vtd::geom;

struct HhComp {
  int32_t start;        // raster index of first pixel
  int32_t external;     // 1 if RETR_EXTERNAL would return it
  int64_t area2;        // twice the signed contour area
  float rect[5];        // cx, cy, w, h, angle
  float box[8];         // boxPoints
  int32_t nhull;
  int32_t pad;
};

extern "C" int hh_components(const uint8_t* mask, int h, int w, HhComp* out, int cap) {
  const int n = h * w;
  std::vector<int> lab(n, -1);
  std::vector<uint8_t> outside(n, 0);
  std::vector<int> stack;
  // background: 4-connected flood from every border background pixel => "outside" region
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (!(y == 0 || x == 0 || y == h - 1 || x == w - 1)) continue;
      int i = y * w + x;
      if (mask[i] || outside[i]) continue;
      outside[i] = 1; stack.push_back(i);
      while (!stack.empty()) {
        int j = stack.back(); stack.pop_back();
        int jy = j / w, jx = j % w;
        const int nx[4] = {jx - 1, jx + 1, jx, jx}, ny[4] = {jy, jy, jy - 1, jy + 1};
        for (int k = 0; k < 4; ++k) {
          if (nx[k] < 0 || ny[k] < 0 || nx[k] >= w || ny[k] >= h) continue;
          int q = ny[k] * w + nx[k];
          if (!mask[q] && !outside[q]) { outside[q] = 1; stack.push_back(q); }
        }
      }
    }
  int count = 0;
  auto fg = [&](int x, int y) { return x >= 0 && y >= 0 && x < w && y < h && mask[y * w + x] != 0; };
  std::vector<int> rowmin, rowmax;
  std::vector<Pt> hull;
  std::vector<float> scratch;
  for (int i = 0; i < n; ++i) {
    if (!mask[i] || lab[i] >= 0) continue;
    // flood the 8-connected component; i is its raster-first pixel
    int ymin = i / w, ymax = ymin;
    lab[i] = i; stack.push_back(i);
    std::vector<int> pix;
    while (!stack.empty()) {
      int j = stack.back(); stack.pop_back(); pix.push_back(j);
      int jy = j / w, jx = j % w;
      ymax = std::max(ymax, jy);
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          int qx = jx + dx, qy = jy + dy;
          if (qx < 0 || qy < 0 || qx >= w || qy >= h) continue;
          int q = qy * w + qx;
          if (mask[q] && lab[q] < 0) { lab[q] = i; stack.push_back(q); }
        }
    }
    if (count >= cap) return -1;
    HhComp& c = out[count++];
    std::memset(&c, 0, sizeof(c));
    c.start = i;
    int x0 = i % w, y0 = i / w;
    c.external = (x0 == 0 || y0 == 0) ? 1 : (outside[i - 1] ? 1 : 0);
    int nrows = ymax - ymin + 1;
    // per-row extremes collected from the border trace itself, as the box extraction kernel does (every row extreme of an
    // 8-connected component lies on its outer border); cross-checked against the flood fill's extremes
    rowmin.assign(nrows, 1 << 30); rowmax.assign(nrows, -1);
    c.area2 = trace_outer_visit(fg, x0, y0, 8LL * n, nullptr, [&](int vx, int vy) {
      rowmin[vy - ymin] = std::min(rowmin[vy - ymin], vx); rowmax[vy - ymin] = std::max(rowmax[vy - ymin], vx);
    });
    for (int j : pix) {
      int jy = j / w - ymin, jx = j % w;
      if (jx < rowmin[jy] || jx > rowmax[jy]) return -2;        // a component pixel outside the traced extremes: cannot happen
    }
    hull.resize(2 * nrows + 2);
    int nh = hull_from_rows(rowmin.data(), rowmax.data(), ymin, nrows, hull.data());
    c.nhull = nh;
    if (nh >= 3) {
      scratch.resize(3 * nh);
      RotRect rr = min_area_rect(hull.data(), nh, scratch.data(), scratch.data() + nh, scratch.data() + 2 * nh);
      c.rect[0] = rr.cx; c.rect[1] = rr.cy; c.rect[2] = rr.w; c.rect[3] = rr.h; c.rect[4] = rr.angle;
      PtF bp[4];
      box_points(rr, bp);
      for (int k = 0; k < 4; ++k) { c.box[2 * k] = bp[k].x; c.box[2 * k + 1] = bp[k].y; }
    }
  }
  return count;
}

// unclip + boxPoints of a min-area rect as the product computes them (box_geom.cuh unclip_rect / box_points)
extern "C" void hh_unclip_box(const float* rect5, float ratio, float* out_rect5, float* out_box8) {
  RotRect rr = {rect5[0], rect5[1], rect5[2], rect5[3], rect5[4]};
  unclip_rect(rr, ratio);
  out_rect5[0] = rr.cx; out_rect5[1] = rr.cy; out_rect5[2] = rr.w; out_rect5[3] = rr.h; out_rect5[4] = rr.angle;
  PtF bp[4];
  box_points(rr, bp);
  for (int k = 0; k < 4; ++k) { out_box8[2 * k] = bp[k].x; out_box8[2 * k + 1] = bp[k].y; }
}

// hull of an arbitrary point list given as row extremes (for direct comparison with cv2.convexHull)
extern "C" int hh_hull_rows(const int* rowmin, const int* rowmax, int y0, int nrows, int* out_xy) {
  std::vector<Pt> hull(2 * nrows + 2);
  int nh = hull_from_rows(rowmin, rowmax, y0, nrows, hull.data());
  for (int i = 0; i < nh; ++i) { out_xy[2 * i] = hull[i].x; out_xy[2 * i + 1] = hull[i].y; }
  return nh;
}


// The product's Pillow coefficient tables (csrc/resize_tab.h, what vtd_preprocess uploads for the resize kernel).
// lo/cnt: [out_size]; kk: [out_size * ksize_cap] (rows padded with zeros); returns ksize, or -1 if ksize > ksize_cap.
extern "C" int hh_resize_tab(int in_size, int out_size, int* lo, int* cnt, int* kk, int ksize_cap, int* maxcnt) {
  std::vector<int> l, c, k;
  int ksize = 0;
  vtd::compute_resize_tab(in_size, out_size, &l, &c, &k, &ksize, maxcnt);
  if (ksize > ksize_cap) return -1;
  for (int i = 0; i < out_size; ++i) {
    lo[i] = l[i]; cnt[i] = c[i];
    for (int x = 0; x < ksize_cap; ++x) kk[(size_t)i * ksize_cap + x] = x < ksize ? k[(size_t)i * ksize + x] : 0;
  }
  return ksize;
}
