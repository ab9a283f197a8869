// Small HBM-bound helpers: max-pooling over NHWC, NCHW<->NHWC conversion for the DBNet.forward /
// CRNN.forward drop-ins, and the strict `prob > thr` binarisation (text_detector.py:144).
#include "common.cuh"
#include <float.h>

namespace vtd {
namespace {

// nn.MaxPool2d semantics (implicit -inf padding). One thread per (pixel, 4-channel group).
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int Ho,
                               int Wo, int kh, int kw, int sh, int sw, int ph, int pw) {
  const int C4 = C >> 2;
  long long total = (long long)N * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    long long p = i / C4;
    int ox = (int)(p % Wo); p /= Wo;
    int oy = (int)(p % Ho);
    int n = (int)(p / Ho);
    float m[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int r = 0; r < kh; ++r) {
      int iy = oy * sh - ph + r;
      if ((unsigned)iy >= (unsigned)H) continue;
      for (int s = 0; s < kw; ++s) {
        int ix = ox * sw - pw + s;
        if ((unsigned)ix >= (unsigned)W) continue;
        const T* q = in + (((long long)n * H + iy) * W + ix) * C + c4 * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = fmaxf(m[j], to_f(q[j]));
      }
    }
    T* o = out + (((long long)n * Ho + oy) * Wo + ox) * C + c4 * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = from_f<T>(m[j]);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int C, int H, int W,
                                    OutLayout lay) {
  const int Cpad = lay.cpp;
  long long total = (long long)N * H * W * Cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cpad);
    long long p = i / Cpad;
    int x = (int)(p % W); p /= W;
    int y = (int)(p % H);
    int n = (int)(p / H);
    float v = c < C ? in[(((long long)n * C + c) * H + y) * W + x] : 0.f;
    out[lay.offset + n * lay.img_pitch + y * lay.row_pitch + (long long)x * Cpad + c] = from_f<T>(v);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W,
                                    OutLayout lay) {
  long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    long long p = i / W;
    int y = (int)(p % H); p /= H;
    int c = (int)(p % C);
    int n = (int)(p / C);
    out[i] = to_f(in[lay.offset + n * lay.img_pitch + y * lay.row_pitch + (long long)x * lay.cpp + c]);
  }
}

__global__ void threshold_kernel(const float* __restrict__ prob, uint8_t* __restrict__ mask, long long count,
                                 float thr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    mask[i] = prob[i] > thr ? 1 : 0;
}

inline int grid_for(long long total, int bs) {
  long long g = (total + bs - 1) / bs;
  long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

template <typename T>
cudaError_t maxpool_nhwc(const T* in, T* out, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph,
                         int pw, cudaStream_t s, LaunchCounter* lc) {
  int Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  long long total = (long long)N * Ho * Wo * (C / 4);
  if (total <= 0) return cudaSuccess;
  maxpool_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(in, out, N, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template <typename T>
cudaError_t nchw_f32_to_nhwc(const float* in, T* out, int N, int C, int H, int W, OutLayout lay, cudaStream_t s,
                             LaunchCounter* lc) {
  long long total = (long long)N * H * W * lay.cpp;
  if (total <= 0) return cudaSuccess;
  nchw_to_nhwc_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(in, out, N, C, H, W, lay);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template <typename T>
cudaError_t nhwc_to_nchw_f32(const T* in, float* out, int N, int C, int H, int W, OutLayout lay, cudaStream_t s,
                             LaunchCounter* lc) {
  long long total = (long long)N * C * H * W;
  if (total <= 0) return cudaSuccess;
  nhwc_to_nchw_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(in, out, N, C, H, W, lay);
  if (lc) lc->n++;
  return cudaGetLastError();
}

cudaError_t threshold_mask(const float* prob, uint8_t* mask, long long count, float thr, cudaStream_t s,
                           LaunchCounter* lc) {
  if (count <= 0) return cudaSuccess;
  threshold_kernel<<<grid_for(count, 256), 256, 0, s>>>(prob, mask, count, thr);
  if (lc) lc->n++;
  return cudaGetLastError();
}

#define INST(T)                                                                                                  \
  template cudaError_t maxpool_nhwc<T>(const T*, T*, int, int, int, int, int, int, int, int, int, int,           \
                                       cudaStream_t, LaunchCounter*);                                            \
  template cudaError_t nchw_f32_to_nhwc<T>(const float*, T*, int, int, int, int, OutLayout, cudaStream_t,        \
                                           LaunchCounter*);                                                      \
  template cudaError_t nhwc_to_nchw_f32<T>(const T*, float*, int, int, int, int, OutLayout, cudaStream_t,        \
                                           LaunchCounter*);
INST(float)
INST(bf16)

}  // namespace vtd
