// Small HBM-bound helpers: max-pooling over NHWC, NCHW<->NHWC conversion for the DBNet.forward /
// CRNN.forward drop-ins, and the strict `prob > thr` binarisation (text_detector.py:144).
#include "common.cuh"
#include <float.h>

namespace vtd {
namespace {

// nn.MaxPool2d semantics (implicit -inf padding).  One thread per (output pixel, 16-byte channel group): all taps
// are requested first (KH*KW independent 16-byte loads in flight), then reduced -- HBM/L2-bound, no reuse needed
// beyond what L2 provides.
template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ uint4 ninf() {
    uint32_t u = 0xff800000u; return make_uint4(u, u, u, u);
  }
  static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) {
    uint4 r;
    r.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)));
    r.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y)));
    r.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)));
    r.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w)));
    return r;
  }
};
template <> struct PoolVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ uint4 ninf() {
    uint32_t u = 0xff80ff80u; return make_uint4(u, u, u, u);
  }
  static __device__ __forceinline__ uint32_t m2(uint32_t a, uint32_t b) {
    bf16x2 r = __hmax2(*reinterpret_cast<bf16x2*>(&a), *reinterpret_cast<bf16x2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) {
    return make_uint4(m2(a.x, b.x), m2(a.y, b.y), m2(a.z, b.z), m2(a.w, b.w));
  }
};

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(256) maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H,
                                                      int W, int C, int Ho, int Wo, int sh, int sw, int ph, int pw) {
  constexpr int V = PoolVec<T>::N;
  const int CV = C / V;
  const long long total = (long long)N * Ho * Wo * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    uint4 t[KH * KW];
#pragma unroll
    for (int r = 0; r < KH; ++r) {
#pragma unroll
      for (int s = 0; s < KW; ++s) {
        const int iy = oy * sh - ph + r, ix = ox * sw - pw + s;
        const bool ok = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
        t[r * KW + s] = ok ? __ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + iy) * W + ix) * C) + cv)
                           : PoolVec<T>::ninf();
      }
    }
    uint4 m = t[0];
#pragma unroll
    for (int k = 1; k < KH * KW; ++k) m = PoolVec<T>::vmax(m, t[k]);
    reinterpret_cast<uint4*>(out + (((long long)n * Ho + oy) * Wo + ox) * C)[cv] = m;
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
  constexpr int V = PoolVec<T>::N;
  if (C % V != 0) return cudaErrorInvalidValue;
  long long total = (long long)N * Ho * Wo * (C / V);
  if (total <= 0) return cudaSuccess;
  long long g = (total + 255) / 256;
  const long long cap = 148LL * 32;
  const int grid = (int)(g < cap ? g : cap);
  if (kh == 3 && kw == 3) maxpool_kernel<T, 3, 3><<<grid, 256, 0, s>>>(in, out, N, H, W, C, Ho, Wo, sh, sw, ph, pw);
  else if (kh == 2 && kw == 2) maxpool_kernel<T, 2, 2><<<grid, 256, 0, s>>>(in, out, N, H, W, C, Ho, Wo, sh, sw, ph, pw);
  else if (kh == 2 && kw == 1) maxpool_kernel<T, 2, 1><<<grid, 256, 0, s>>>(in, out, N, H, W, C, Ho, Wo, sh, sw, ph, pw);
  else return cudaErrorInvalidValue;
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
