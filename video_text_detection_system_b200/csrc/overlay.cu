// Annotated-frame overlay on the device: ProcessingService._draw_detections
// (app/services/processing_service.py:188-218) -- per detection, in list order,
//     cv2.rectangle(frame, (x1, y1), (x2, y2), (0, 255, 0), 2)
//     (tw, th) = cv2.getTextSize(label, FONT_HERSHEY_SIMPLEX, 0.5, 1)[0]
//     cv2.rectangle(frame, (x1, y1 - th - 10), (x1 + tw, y1), (0, 255, 0), -1)
//     cv2.putText(frame, label, (x1, y1 - 5), FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)
// on BGR uint8 frames that are (or were just copied) in HBM.
//
// A later primitive overwrites an earlier one, so the kernel is a gather, not a scatter: one CTA per detection walks the
// pixels of that detection's footprint and, for each, finds the LAST primitive of the frame's draw list that covers it (its
// own three, then those of the later detections whose footprints intersect -- normally none); only then does it store the
// colour.  Two CTAs that reach the same pixel compute the same answer, so duplicate stores are benign and no pass over whole
// frames, atomics or ordering between CTAs is needed.
//
// What "covers" means, matched to OpenCV 4.x pixel for pixel:
//   * thickness-2 outline of an axis-aligned rectangle = the 3-pixel bands centred on its four edges, minus the four outer
//     corner pixels (ThickLine's rounded ends);
//   * filled plate = the inclusive rectangle between its two corners;
//   * text = glyph cells from overlay_atlas.h (rendered by OpenCV itself, one cell per byte and sub-pixel pen phase), pen
//     advancing OV_WIDTHS[c] half-pixels per glyph; bytes outside 32..126 draw '?' (putText's readCheck);
//   * getTextSize: tw = rint(sum of widths / 2 + 1) (cvRound = round-half-even), th = OV_TEXT_H.
//   * a glyph cut by ONE frame border: OpenCV clips a stroke that crosses the border in 16.16 fixed point before rasterising
//     it, which can move one or two of the stroke's remaining pixels; overlay_atlas.h holds the 1 161 (glyph, phase, border,
//     distance) cells where that happens, found by key (binary search), all others are the plain cell cropped.
// One known difference is left, measured in tests/test_gpu_overlay.py: a glyph cut by TWO borders at once (a label in a frame
// corner) is cropped from the plain cell.
#include "common.cuh"
#include "overlay_atlas.h"

namespace vtd {
namespace {

struct OvConst {
  uint16_t cells[OV_LAST - OV_FIRST + 1][2][OV_CELL_H];
  uint8_t widths[OV_LAST - OV_FIRST + 1];
};
__constant__ OvConst ov_tab;
__device__ uint32_t ov_patch_keys[OV_PATCHES];
__device__ uint16_t ov_patch_cells[OV_PATCHES][OV_CELL_H];

// the cell row of glyph g at pen phase ph, row `row`, when the cell's top-left pixel is (px, ty0) in an h x w frame
__device__ uint32_t glyph_row(int g, int ph, int row, int px, int ty0, int h, int w) {
  const int kt = ty0 < 0 ? -ty0 : 0, kb = ty0 + OV_CELL_H - 1 > h - 1 ? ty0 + OV_CELL_H - h : 0;
  const int kl = px < 0 ? -px : 0, kr = px + OV_CELL_W - 1 > w - 1 ? px + OV_CELL_W - w : 0;
  const int sides = (kt > 0) + (kb > 0) + (kl > 0) + (kr > 0);
  if (sides == 1) {
    const int side = kt ? 0 : (kb ? 1 : (kl ? 2 : 3)), k = kt + kb + kl + kr;
    if (k <= 16) {
      const uint32_t key = (uint32_t)(((g * 2 + ph) * 4 + side) * 17 + k);
      int lo = 0, hi = OV_PATCHES - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t km = ov_patch_keys[mid];
        if (km == key) return ov_patch_cells[mid][row];
        if (km < key) lo = mid + 1; else hi = mid - 1;
      }
    }
  }
  return ov_tab.cells[g][ph][row];
}

struct Foot {            // inclusive pixel ranges, unclipped
  int ax, ay, bx, by;    // rectangle corners, ordered
  int px0, py0, px1, py1;  // plate
  int tx0, tx1, ty0, ty1;  // text cells
  int x0, y0, x1, y1;    // everything
};

__device__ __forceinline__ int glyph_of(uint8_t b) { return (b < OV_FIRST || b > OV_LAST) ? '?' - OV_FIRST : b - OV_FIRST; }

__device__ Foot footprint(const OverlayItem& it, int sumw) {
  Foot f;
  const int x1 = it.bbox[0], y1 = it.bbox[1], x2 = it.bbox[2], y2 = it.bbox[3];
  f.ax = min(x1, x2); f.bx = max(x1, x2); f.ay = min(y1, y2); f.by = max(y1, y2);
  const int tw = (int)rint(sumw * 0.5 + 1.0);
  f.px0 = x1; f.px1 = x1 + tw; f.py0 = y1 - OV_TEXT_H - 10; f.py1 = y1;
  f.tx0 = x1; f.tx1 = x1 + (sumw >> 1) + OV_CELL_W; f.ty0 = y1 - 5 + OV_ROW0; f.ty1 = f.ty0 + OV_CELL_H - 1;
  f.x0 = min(f.ax - 1, f.px0); f.x1 = max(max(f.bx + 1, f.px1), f.tx1);
  f.y0 = min(f.ay - 1, f.py0); f.y1 = max(f.by + 1, f.py1);
  return f;
}

__device__ __forceinline__ int label_width(const OverlayItem& it) {
  int s = 0;
  for (int k = 0; k < it.label_len; ++k) s += ov_tab.widths[glyph_of(it.label[k])];
  return s;
}

// pens: half-pixel pen position before glyph k (prefix sums), or nullptr to accumulate on the fly
__device__ bool text_covers(const OverlayItem& it, const Foot& f, const int* pens, int x, int y, int h, int w) {
  if (x < f.tx0 || x > f.tx1 || y < f.ty0 || y > f.ty1) return false;
  const int row = y - f.ty0;
  const bool inside = f.ty0 >= 0 && f.ty1 < h && f.tx0 >= 0 && f.tx1 < w;       // the whole label: no glyph is cut
  int pen2 = 2 * it.bbox[0];
  for (int k = 0; k < it.label_len; ++k) {
    const int g = glyph_of(it.label[k]);
    if (pens) pen2 = pens[k];
    const int b = x - (pen2 >> 1);
    if (b >= 0 && b < OV_CELL_W) {
      const uint32_t bits = inside ? ov_tab.cells[g][pen2 & 1][row] : glyph_row(g, pen2 & 1, row, pen2 >> 1, f.ty0, h, w);
      if ((bits >> b) & 1) return true;
    }
    if (!pens) pen2 += ov_tab.widths[g];
  }
  return false;
}

// 0 = untouched, 1 = green, 2 = black: the last of the detection's three primitives that covers (x, y)
__device__ int item_covers(const OverlayItem& it, const Foot& f, const int* pens, int x, int y, int h, int w) {
  if (x < f.x0 || x > f.x1 || y < f.y0 || y > f.y1) return 0;
  if (text_covers(it, f, pens, x, y, h, w)) return 2;
  if (x >= f.px0 && x <= f.px1 && y >= f.py0 && y <= f.py1) return 1;
  if (x >= f.ax - 1 && x <= f.bx + 1 && y >= f.ay - 1 && y <= f.by + 1) {
    const bool band = x <= f.ax + 1 || x >= f.bx - 1 || y <= f.ay + 1 || y >= f.by - 1;
    const bool corner = (x == f.ax - 1 || x == f.bx + 1) && (y == f.ay - 1 || y == f.by + 1);
    if (band && !corner) return 1;
  }
  return 0;
}

constexpr int OV_THREADS = 256;
constexpr int OV_MAX_LATER = 256;

// items are grouped by frame (frame_end[i] = one past the last item of item i's frame), draw order inside a frame
__global__ void __launch_bounds__(OV_THREADS) overlay_kernel(uint8_t* const* __restrict__ frames, int h, int w, int pitch,
                                                             const OverlayItem* __restrict__ items,
                                                             const int* __restrict__ frame_end) {
  __shared__ int pens[OV_LABEL_MAX];
  __shared__ Foot me;
  __shared__ int later[OV_MAX_LATER];
  __shared__ Foot later_foot[8];
  __shared__ int n_later;
  const int i = blockIdx.x;
  const OverlayItem& it = items[i];
  if (threadIdx.x == 0) {
    int pen2 = 2 * it.bbox[0];
    for (int k = 0; k < it.label_len; ++k) { pens[k] = pen2; pen2 += ov_tab.widths[glyph_of(it.label[k])]; }
    me = footprint(it, pen2 - 2 * it.bbox[0]);
    n_later = 0;
  }
  __syncthreads();
  const int end = frame_end[i];
  for (int j = i + 1 + threadIdx.x; j < end; j += OV_THREADS) {
    const Foot fj = footprint(items[j], label_width(items[j]));
    if (fj.x0 <= me.x1 && fj.x1 >= me.x0 && fj.y0 <= me.y1 && fj.y1 >= me.y0) {
      const int slot = atomicAdd(&n_later, 1);
      if (slot < OV_MAX_LATER) later[slot] = j;
    }
  }
  __syncthreads();
  const int nl = min(n_later, OV_MAX_LATER);
  // the atomics filled `later` in no particular order; what matters per pixel is the largest covering index
  if (threadIdx.x < 8 && threadIdx.x < nl) later_foot[threadIdx.x] = footprint(items[later[threadIdx.x]], label_width(items[later[threadIdx.x]]));
  __syncthreads();
  uint8_t* frame = frames[it.frame];
  // five rectangles hold every pixel this detection can touch: the four edge bands and the label (plate + text cells)
  for (int part = 0; part < 5; ++part) {
    int rx0, ry0, rx1, ry1;
    if (part == 0) { rx0 = me.ax - 1; rx1 = me.bx + 1; ry0 = me.ay - 1; ry1 = min(me.ay + 1, me.by + 1); }
    else if (part == 1) { rx0 = me.ax - 1; rx1 = me.bx + 1; ry0 = max(me.by - 1, me.ay + 2); ry1 = me.by + 1; }
    else if (part == 2) { rx0 = me.ax - 1; rx1 = min(me.ax + 1, me.bx + 1); ry0 = me.ay + 2; ry1 = me.by - 2; }
    else if (part == 3) { rx0 = max(me.bx - 1, me.ax + 2); rx1 = me.bx + 1; ry0 = me.ay + 2; ry1 = me.by - 2; }
    else { rx0 = me.px0; rx1 = max(me.px1, me.tx1); ry0 = me.py0; ry1 = me.py1; }
    rx0 = max(rx0, 0); ry0 = max(ry0, 0); rx1 = min(rx1, w - 1); ry1 = min(ry1, h - 1);
    if (rx0 > rx1 || ry0 > ry1) continue;
    const int rw = rx1 - rx0 + 1;
    const long long area = (long long)rw * (ry1 - ry0 + 1);
    for (long long q = threadIdx.x; q < area; q += OV_THREADS) {
      const int x = rx0 + (int)(q % rw), y = ry0 + (int)(q / rw);
      int colour = item_covers(it, me, pens, x, y, h, w), best = colour ? i : -1;
      for (int s = 0; s < nl; ++s) {
        const int j = later[s];
        if (j < best) continue;
        Foot fj;
        if (s < 8) fj = later_foot[s]; else fj = footprint(items[j], label_width(items[j]));
        const int cj = item_covers(items[j], fj, nullptr, x, y, h, w);
        if (cj) { colour = cj; best = j; }
      }
      if (colour) {
        uint8_t* px = frame + (size_t)y * pitch + (size_t)x * 3;
        px[0] = 0; px[1] = colour == 1 ? 255 : 0; px[2] = 0;
      }
    }
  }
}

}  // namespace

cudaError_t overlay_upload_tables(cudaStream_t s) {
  static OvConst host;
  memcpy(host.cells, OV_CELLS, sizeof(host.cells));
  memcpy(host.widths, OV_WIDTHS, sizeof(host.widths));
  cudaError_t e = cudaMemcpyToSymbolAsync(ov_tab, &host, sizeof(host), 0, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(ov_patch_keys, OV_PATCH_KEYS, sizeof(OV_PATCH_KEYS), 0, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(ov_patch_cells, OV_PATCH_CELLS, sizeof(OV_PATCH_CELLS), 0, cudaMemcpyHostToDevice, s);
  return e;
}

cudaError_t draw_overlay(uint8_t* const* frames, int h, int w, int pitch, const OverlayItem* items, const int* frame_end,
                         int n_items, cudaStream_t s, LaunchCounter* lc) {
  if (n_items <= 0) return cudaSuccess;
  overlay_kernel<<<n_items, OV_THREADS, 0, s>>>(frames, h, w, pitch, items, frame_end);
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
