// Pillow resize coefficient tables (ImagingResample.c: precompute_coeffs + normalize_coeffs_8bpc) for the BILINEAR
// filter: support 1.0, widened by the scale when down-sampling (antialias), weights normalised in double and rounded to
// 22-bit fixed point.  Replaces what transforms.Resize -> PIL.Image.resize computes per call (text_detector.py:101).
// Host-only and free of CUDA so that tests/host_harness.cpp can sweep it against the oracle on the CPU.
#pragma once
#include <cmath>
#include <vector>

namespace vtd {

// lo[xx], cnt[xx]: first source index and tap count of output xx; kk[xx*ksize + x]: its taps (zero padded to ksize).
inline void compute_resize_tab(int in_size, int out_size, std::vector<int>* lo_out, std::vector<int>* cnt_out,
                               std::vector<int>* kk_out, int* ksize_out, int* maxcnt_out) {
  const double scale = (double)in_size / (double)out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * fs;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  std::vector<int> lo(out_size), cnt(out_size), kk((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  const double ss = 1.0 / fs;
  int maxcnt = 0;
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5); if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5); if (xmax > in_size) xmax = in_size;
    int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      double v = a < 1.0 ? 1.0 - a : 0.0;
      w[x] = v; ww += v;
    }
    for (int x = 0; x < n; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (double)(1 << 22)) : (int)(0.5 + v * (double)(1 << 22));
    }
    lo[xx] = xmin; cnt[xx] = n;
    if (n > maxcnt) maxcnt = n;
  }
  lo_out->swap(lo); cnt_out->swap(cnt); kk_out->swap(kk);
  *ksize_out = ksize; *maxcnt_out = maxcnt;
}

}  // namespace vtd
