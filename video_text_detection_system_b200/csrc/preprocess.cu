// Stage 1: frame preprocessing.
//
// Replaces, bit for bit, what TextDetector.detect does on the CPU before the network
// (text_detector.py:99-104,117-124): cv2.cvtColor(BGR2RGB) -> ToPILImage -> Resize((Hd,Wd))
// [= PIL.Image.resize, BILINEAR, antialiased when shrinking] -> ToTensor (/255) -> Normalize.
//
// Pillow's resample (ImagingResample) is separable: a horizontal pass into a uint8 intermediate,
// then a vertical pass; each pass is a dot product of `cnt` source samples with 22-bit fixed-point
// weights, (2^21 + sum) >> 22, clipped to [0,255].  The host builds the per-axis tables
// (lo, cnt, kk) once per (source size, detector size) pair; this kernel applies them.
//
// One CTA produces a TH x TW tile of the output: it runs the horizontal pass for the source rows
// the tile needs straight from global memory (neighbouring lanes read neighbouring bytes, L1 serves
// the overlap), keeps the uint8 intermediate in shared memory, runs the vertical pass from there,
// normalises and writes NHWC with C padded to 4 (one 8/16-byte vector store per pixel).  The
// intermediate never touches HBM.  HBM-bound: algorithmic bytes = h*w*3 read + Hd*Wd*3*sizeof(T) written.
#include "common.cuh"

namespace vtd {
namespace {

constexpr int TH = 16, TW = 64, NT = 256;

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// ITU-R BT.601 limited-range YUV -> RGB in 20-bit fixed point (the arithmetic cv2.cvtColor uses for
// COLOR_YUV2BGR_NV12).  Returns channel c of (B,G,R).
__device__ __forceinline__ int nv12_bgr(const uint8_t* f, int h, int pitch, int y, int x, int c) {
  int Y = f[(size_t)y * pitch + x];
  const uint8_t* uv = f + (size_t)h * pitch + (size_t)(y >> 1) * pitch + (x & ~1);
  int u = (int)uv[0] - 128, v = (int)uv[1] - 128;
  int yy = max(0, Y - 16) * 1220542;
  int val;
  if (c == 2) val = (yy + (1 << 19) + 1673527 * v) >> 20;
  else if (c == 1) val = (yy + (1 << 19) - 852492 * v - 409993 * u) >> 20;
  else val = (yy + (1 << 19) + 2116026 * u) >> 20;
  return clip8(val);
}

template <typename T, int PIX>
__global__ void __launch_bounds__(NT) preprocess_kernel(const uint8_t* const* __restrict__ frames, int h, int w,
                                                        int pitch, ResizeTab tx, ResizeTab ty, int rows_cap,
                                                        T* __restrict__ out, OutLayout lay) {
  extern __shared__ uint8_t tmp[];   // [rows][TW][3] horizontal-pass result, uint8 like Pillow's
  const uint8_t* __restrict__ f = frames[blockIdx.z];
  const int dw = tx.out_size, dh = ty.out_size;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const int oy1 = min(oy0 + TH, dh) - 1;
  const int r0 = ty.lo[oy0];
  const int r1 = ty.lo[oy1] + ty.cnt[oy1];       // exclusive (lo is non-decreasing in oy)
  const int rows = min(r1 - r0, rows_cap);
  const int tw = min(TW, dw - ox0);

  // ---- horizontal pass: items = rows x tw x 3
  for (int it = threadIdx.x; it < rows * tw * 3; it += NT) {
    int c = it % 3;
    int q = it / 3;
    int xo = q % tw, r = q / tw;
    int ox = ox0 + xo, iy = r0 + r;
    int lo = tx.lo[ox], cnt = tx.cnt[ox];
    const int* kk = tx.kk + (size_t)ox * tx.ksize;
    int acc = 1 << 21;
    if (PIX == 0) {
      const uint8_t* p = f + (size_t)iy * pitch + (size_t)lo * 3 + c;
      for (int j = 0; j < cnt; ++j) acc += (int)p[j * 3] * kk[j];
    } else {
      for (int j = 0; j < cnt; ++j) acc += nv12_bgr(f, h, pitch, iy, lo + j, c) * kk[j];
    }
    tmp[(r * TW + xo) * 3 + c] = (uint8_t)clip8(acc >> 22);
  }
  __syncthreads();

  // ---- vertical pass + normalise: items = th x tw pixels
  const int th = oy1 - oy0 + 1;
  for (int it = threadIdx.x; it < th * tw; it += NT) {
    int xo = it % tw, yo = it / tw;
    int oy = oy0 + yo;
    int lo = ty.lo[oy] - r0, cnt = ty.cnt[oy];
    const int* kk = ty.kk + (size_t)oy * ty.ksize;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int j = 0; j < cnt; ++j) {
      const uint8_t* p = tmp + ((lo + j) * TW + xo) * 3;
      int k = kk[j];
      a0 += (int)p[0] * k; a1 += (int)p[1] * k; a2 += (int)p[2] * k;
    }
    // source order is B,G,R; the network wants R,G,B (cvtColor at text_detector.py:120)
    float b = (float)clip8(a0 >> 22), g = (float)clip8(a1 >> 22), r = (float)clip8(a2 >> 22);
    // ToTensor: x/255 (fp32 divide); Normalize: (x-mean)/std (fp32 subtract, fp32 divide) -- IEEE, no fast-math
    float v[4];
    v[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.0f), 0.485f), 0.229f);
    v[1] = __fdiv_rn(__fsub_rn(__fdiv_rn(g, 255.0f), 0.456f), 0.224f);
    v[2] = __fdiv_rn(__fsub_rn(__fdiv_rn(b, 255.0f), 0.406f), 0.225f);
    v[3] = 0.f;
    size_t o = (size_t)(lay.offset + blockIdx.z * lay.img_pitch + oy * lay.row_pitch + (long long)(ox0 + xo) * 4);
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + o) = u;
    }
  }
}

}  // namespace

template <typename T>
cudaError_t preprocess_frames(const uint8_t* const* frames_dev, int n, int h, int w, int pitch, int pixfmt,
                              const ResizeTab& tx, const ResizeTab& ty, uint8_t* /*tmp_u8*/, T* out, OutLayout lay,
                              cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  // rows of the intermediate one tile can need: TH output rows span at most TH*scale + ksize source rows
  double scale = (double)ty.in_size / ty.out_size;
  int rows_cap = (int)(TH * (scale > 1.0 ? scale : 1.0)) + ty.ksize + 2;
  size_t smem = (size_t)rows_cap * TW * 3;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  dim3 grid((tx.out_size + TW - 1) / TW, (ty.out_size + TH - 1) / TH, n);
  cudaError_t e;
  if (pixfmt == 0) {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(preprocess_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    preprocess_kernel<T, 0><<<grid, NT, smem, s>>>(frames_dev, h, w, pitch, tx, ty, rows_cap, out, lay);
  } else {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(preprocess_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    preprocess_kernel<T, 1><<<grid, NT, smem, s>>>(frames_dev, h, w, pitch, tx, ty, rows_cap, out, lay);
  }
  if (lc) lc->n++;
  return cudaGetLastError();
}

template cudaError_t preprocess_frames<float>(const uint8_t* const*, int, int, int, int, int, const ResizeTab&,
                                              const ResizeTab&, uint8_t*, float*, OutLayout, cudaStream_t, LaunchCounter*);
template cudaError_t preprocess_frames<bf16>(const uint8_t* const*, int, int, int, int, int, const ResizeTab&,
                                             const ResizeTab&, uint8_t*, bf16*, OutLayout, cudaStream_t, LaunchCounter*);

}  // namespace vtd
