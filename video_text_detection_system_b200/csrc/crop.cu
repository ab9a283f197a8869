// Stage 4b (first half): batched crop + resize gather into CRNN inputs.
//
// Replaces, for every detection of a batch at once, `crop = frame[y1:y2, x1:x2]` (pipeliine.py:121) and
// `cv2.resize(img, (128, 32))` + HWC->CHW + `/255` (text_recognizer.py:118-119).  The crop is taken from the
// ORIGINAL BGR frame (channel order kept, no mean/std).  cv2.resize INTER_LINEAR on uint8 is fixed point:
// half-pixel centres, 11-bit coefficients, horizontal pass to int32, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2  (oracle/port.py
// cv_resize_linear_restated).  Output is NHWC with C padded to 4, in the activation type.
//
// One CTA per crop; boxes are read from the vtd_record array the box-extraction stage left in device
// memory, so there is no host round trip between detection and recognition.  HBM-bound gather:
// algorithmic bytes = source box area * 3 read + 32*crop_w*4*sizeof(T) written per crop.
#include "common.cuh"
#include "../../include/vtd.h"

namespace vtd {
namespace {

constexpr int CH = 32;    // crop height (text_recognizer.py:118)

struct Tap { int s0, s1, a0, a1; };

// cv::resize linear coefficient for destination index d (n_in -> n_out)
__device__ __forceinline__ Tap cv_tap(int d, int n_in, int n_out) {
  double sc = (double)n_in / (double)n_out;
  float f = (float)(((double)d + 0.5) * sc - 0.5);
  int s = (int)floorf(f);
  f = f - (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= n_in - 1) { s = n_in - 1; f = 0.f; }
  Tap t;
  t.s0 = s;
  t.s1 = min(s + 1, n_in - 1);
  t.a1 = __float2int_rn(f * 2048.f);
  t.a0 = __float2int_rn((1.f - f) * 2048.f);
  return t;
}

// one pixel = cpp (4 or 8) channels: b, g, r, then zeros
template <typename T>
__device__ __forceinline__ void store_px(T* o, float b, float g, float r, int cpp);
template <> __device__ __forceinline__ void store_px<float>(float* o, float b, float g, float r, int cpp) {
  *reinterpret_cast<float4*>(o) = make_float4(b, g, r, 0.f);
  if (cpp == 8) *reinterpret_cast<float4*>(o + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <> __device__ __forceinline__ void store_px<bf16>(bf16* o, float b, float g, float r, int cpp) {
  bf16x2 p0 = pack2(b, g), p1 = pack2(r, 0.f);
  if (cpp == 8) {
    uint4 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1); u.z = 0u; u.w = 0u;
    *reinterpret_cast<uint4*>(o) = u;
  } else {
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(o) = u;
  }
}

// BT.601 limited-range NV12 -> B,G,R (cv2.cvtColor COLOR_YUV2BGR_NV12 fixed point, bit-exact): the decoder-surface
// ingest path crops straight from the NV12 frame, which equals converting the frame first and cropping after.
__device__ __forceinline__ int nv12_px(const uint8_t* __restrict__ f, int frame_h, int pitch, int y, int x, int c) {
  int Y = f[(size_t)y * pitch + x];
  const uint8_t* uv = f + (size_t)frame_h * pitch + (size_t)(y >> 1) * pitch + (x & ~1);
  int u = (int)uv[0] - 128, v = (int)uv[1] - 128;
  int yy = max(0, Y - 16) * 1220542;
  int val = c == 2 ? (yy + (1 << 19) + 1673527 * v) >> 20
                   : (c == 1 ? (yy + (1 << 19) - 852492 * v - 409993 * u) >> 20 : (yy + (1 << 19) + 2116026 * u) >> 20);
  return val < 0 ? 0 : (val > 255 ? 255 : val);
}

// NV12 = true: `src` is the whole frame (frame_h rows of Y then UV), (x0,y0) the crop origin
template <typename T, bool NV12>
__device__ void resize_one(const uint8_t* __restrict__ src, int pitch, int h, int w, int crop_w,
                           T* __restrict__ out /*pixel (0,0) of this crop*/, OutLayout lay, Tap* xt /*smem [crop_w]*/,
                           Tap* yt /*smem [CH]*/, int frame_h = 0, int x0 = 0, int y0 = 0) {
  for (int i = threadIdx.x; i < crop_w; i += blockDim.x) xt[i] = cv_tap(i, w, crop_w);
  for (int i = threadIdx.x; i < CH; i += blockDim.x) yt[i] = cv_tap(i, h, CH);
  __syncthreads();
  for (int i = threadIdx.x; i < CH * crop_w; i += blockDim.x) {
    const int dx = i % crop_w, dy = i / crop_w;
    const Tap tx = xt[dx], ty = yt[dy];
    const uint8_t* r0 = src + (size_t)ty.s0 * pitch;
    const uint8_t* r1 = src + (size_t)ty.s1 * pitch;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int p00, p01, p10, p11;
      if (NV12) {
        p00 = nv12_px(src, frame_h, pitch, y0 + ty.s0, x0 + tx.s0, c); p01 = nv12_px(src, frame_h, pitch, y0 + ty.s0, x0 + tx.s1, c);
        p10 = nv12_px(src, frame_h, pitch, y0 + ty.s1, x0 + tx.s0, c); p11 = nv12_px(src, frame_h, pitch, y0 + ty.s1, x0 + tx.s1, c);
      } else {
        p00 = r0[tx.s0 * 3 + c]; p01 = r0[tx.s1 * 3 + c]; p10 = r1[tx.s0 * 3 + c]; p11 = r1[tx.s1 * 3 + c];
      }
      int h0 = p00 * tx.a0 + p01 * tx.a1;
      int h1 = p10 * tx.a0 + p11 * tx.a1;
      int o = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o = o < 0 ? 0 : (o > 255 ? 255 : o);
      v[c] = __fdiv_rn((float)o, 255.0f);                      // .float() / 255.0
    }
    store_px<T>(out + dy * lay.row_pitch + (long long)dx * lay.cpp, v[0], v[1], v[2], lay.cpp);
  }
}

// records: [n][kmax]; offsets: [n+1] exclusive prefix of counts; one CTA per crop.
template <typename T, bool NV12>
__global__ void __launch_bounds__(256) crop_records_kernel(const uint8_t* const* __restrict__ frames, int src_h,
                                                           int src_w, int pitch,
                                                           const vtd_record* __restrict__ records,
                                                           const int* __restrict__ offsets, int n, int kmax,
                                                           int first_crop, int crop_w, T* __restrict__ out,
                                                           OutLayout lay) {
  extern __shared__ Tap taps[];
  const int ci = first_crop + blockIdx.x;         // global crop index inside the batch
  if (ci >= offsets[n]) return;
  int f = 0;
  while (f + 1 < n && offsets[f + 1] <= ci) ++f;
  const vtd_record& r = records[(size_t)f * kmax + (ci - offsets[f])];
  // numpy slicing clamps to the frame
  int x1 = min(max(r.bbox[0], 0), src_w), x2 = min(max(r.bbox[2], 0), src_w);
  int y1 = min(max(r.bbox[1], 0), src_h), y2 = min(max(r.bbox[3], 0), src_h);
  int w = x2 - x1, h = y2 - y1;
  T* o = out + lay.offset + blockIdx.x * lay.img_pitch;
  if (w <= 0 || h <= 0) {                          // cannot happen after the >10 size filter; keep the slot defined
    for (int i = threadIdx.x; i < CH * crop_w; i += blockDim.x)
      store_px<T>(o + (i / crop_w) * lay.row_pitch + (long long)(i % crop_w) * lay.cpp, 0.f, 0.f, 0.f, lay.cpp);
    return;
  }
  if (NV12) {
    resize_one<T, true>(frames[f], pitch, h, w, crop_w, o, lay, taps, taps + crop_w, src_h, x1, y1);
  } else {
    const uint8_t* src = frames[f] + (size_t)y1 * pitch + (size_t)x1 * 3;
    resize_one<T, false>(src, pitch, h, w, crop_w, o, lay, taps, taps + crop_w);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) crop_list_kernel(const uint8_t* const* __restrict__ crops,
                                                        const int* __restrict__ hs, const int* __restrict__ ws,
                                                        const int* __restrict__ pitches, int crop_w,
                                                        T* __restrict__ out, OutLayout lay) {
  extern __shared__ Tap taps[];
  const int ci = blockIdx.x;
  resize_one<T, false>(crops[ci], pitches[ci], hs[ci], ws[ci], crop_w, out + lay.offset + ci * lay.img_pitch, lay, taps,
                       taps + crop_w);
}

__global__ void scan_counts_kernel(const int* __restrict__ counts, int n, int* __restrict__ offsets) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < n; ++i) { offsets[i] = acc; acc += counts[i]; }
    offsets[n] = acc;
  }
}

}  // namespace

cudaError_t scan_counts(const int* counts, int n, int* offsets, cudaStream_t s, LaunchCounter* lc) {
  scan_counts_kernel<<<1, 32, 0, s>>>(counts, n, offsets);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template <typename T>
cudaError_t crop_resize_records(const uint8_t* const* frames_dev, int src_h, int src_w, int pitch,
                                const void* records, const int* offsets, int n, int kmax, int first_crop,
                                int n_crops, int crop_w, int nv12, T* out, OutLayout lay, cudaStream_t s,
                                LaunchCounter* lc) {
  if (n_crops <= 0) return cudaSuccess;
  size_t smem = sizeof(Tap) * (crop_w + CH);
  if (nv12)
    crop_records_kernel<T, true><<<n_crops, 256, smem, s>>>(frames_dev, src_h, src_w, pitch,
                                                            reinterpret_cast<const vtd_record*>(records), offsets, n,
                                                            kmax, first_crop, crop_w, out, lay);
  else
    crop_records_kernel<T, false><<<n_crops, 256, smem, s>>>(frames_dev, src_h, src_w, pitch,
                                                             reinterpret_cast<const vtd_record*>(records), offsets, n,
                                                             kmax, first_crop, crop_w, out, lay);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template <typename T>
cudaError_t crop_resize_list(const uint8_t* const* crops_dev, const int* h, const int* w, const int* pitch,
                             int n_crops, int crop_w, T* out, OutLayout lay, cudaStream_t s, LaunchCounter* lc) {
  if (n_crops <= 0) return cudaSuccess;
  size_t smem = sizeof(Tap) * (crop_w + CH);
  crop_list_kernel<T><<<n_crops, 256, smem, s>>>(crops_dev, h, w, pitch, crop_w, out, lay);
  if (lc) lc->n++;
  return cudaGetLastError();
}

#define INST(T)                                                                                                  \
  template cudaError_t crop_resize_records<T>(const uint8_t* const*, int, int, int, const void*, const int*, int, \
                                              int, int, int, int, int, T*, OutLayout, cudaStream_t,              \
                                              LaunchCounter*);                                                   \
  template cudaError_t crop_resize_list<T>(const uint8_t* const*, const int*, const int*, const int*, int, int,  \
                                           T*, OutLayout, cudaStream_t, LaunchCounter*);
INST(float)
INST(bf16)

}  // namespace vtd
