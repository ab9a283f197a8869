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

// Shared-memory plan of one CTA (TH x TW output tile):
//   sx_lo/sx_cnt[TW], sx_k[TW][ksx]   horizontal coefficient slice        (ints)
//   sy_lo/sy_cnt[TH], sy_k[TH][ksy]   vertical coefficient slice
//   src[rows][srcb]                   the source rows the tile needs, BGR bytes, fetched with 16-byte loads
//   tmp[rows][TW][3]                  horizontal-pass result, uint8 like Pillow's intermediate image
template <typename T, int PIX>
__global__ void __launch_bounds__(NT) preprocess_kernel(const uint8_t* const* __restrict__ frames, int h, int w,
                                                        int pitch, ResizeTab tx, ResizeTab ty, int rows_cap,
                                                        int srcb_cap, const float* __restrict__ lut_g,
                                                        T* __restrict__ out, OutLayout lay) {
  extern __shared__ __align__(16) uint8_t smem[];
  // coefficient rows are kept at a pitch of 4 ints (one 16-byte load) when no output sample has more than 4 taps
  const bool fastx = tx.maxcnt <= 4, fasty = ty.maxcnt <= 4;
  const int ksx = fastx ? 4 : tx.ksize, ksy = fasty ? 4 : ty.ksize;
  int* sx_lo = reinterpret_cast<int*>(smem);
  int* sx_cnt = sx_lo + TW;
  int* sx_k = sx_cnt + TW;
  int* sy_lo = sx_k + TW * ksx;
  int* sy_cnt = sy_lo + TH;
  int* sy_k = sy_cnt + TH;
  float* lut = reinterpret_cast<float*>(sy_k + TH * ksy);          // [3][256]: u8 -> (x/255 - mean)/std, IEEE fp32
  uint8_t* src = reinterpret_cast<uint8_t*>(lut + 768);
  uint32_t* tmp = reinterpret_cast<uint32_t*>(src + (size_t)rows_cap * srcb_cap);   // [rows + 4][TW] B | G<<8 | R<<16

  const uint8_t* __restrict__ f = frames[blockIdx.z];
  const int dw = tx.out_size, dh = ty.out_size;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const int tw = min(TW, dw - ox0), th = min(TH, dh - oy0);
  const int tid = threadIdx.x;

  for (int i = tid; i < tw; i += NT) { sx_lo[i] = tx.lo[ox0 + i]; sx_cnt[i] = tx.cnt[ox0 + i]; }
  for (int i = tid; i < tw * ksx; i += NT) {
    const int xo = i / ksx, j = i - xo * ksx;
    sx_k[i] = j < tx.ksize ? tx.kk[(size_t)(ox0 + xo) * tx.ksize + j] : 0;
  }
  for (int i = tid; i < th; i += NT) { sy_lo[i] = ty.lo[oy0 + i]; sy_cnt[i] = ty.cnt[oy0 + i]; }
  for (int i = tid; i < th * ksy; i += NT) {
    const int yo = i / ksy, j = i - yo * ksy;
    sy_k[i] = j < ty.ksize ? ty.kk[(size_t)(oy0 + yo) * ty.ksize + j] : 0;
  }
  for (int i = tid; i < 768; i += NT) lut[i] = __ldg(lut_g + i);
  const int r0 = ty.lo[oy0];
  const int r1 = ty.lo[oy0 + th - 1] + ty.cnt[oy0 + th - 1];       // exclusive (lo is non-decreasing)
  const int rows = min(r1 - r0, rows_cap);
  const int c0 = tx.lo[ox0];
  const int c1 = tx.lo[ox0 + tw - 1] + tx.cnt[ox0 + tw - 1];       // exclusive source column
  const int ncols = c1 - c0;

  // ---- stage the source rows (bytes [c0*3, c1*3) of rows r0..r0+rows) with aligned 16-byte loads
  if (PIX == 0) {
    const int b0 = c0 * 3, b1 = c1 * 3;
    for (int r = tid / 32; r < rows; r += NT / 32) {
      const uint8_t* row = f + (size_t)(r0 + r) * pitch;
      const size_t a0 = (reinterpret_cast<size_t>(row) + b0) & ~(size_t)15;     // aligned start address
      const int skew = (int)(reinterpret_cast<size_t>(row) + b0 - a0);          // bytes before b0 in the first vector
      const int nvec = (skew + (b1 - b0) + 15) >> 4;
      uint8_t* dst = src + (size_t)r * srcb_cap;
      const size_t row_end = reinterpret_cast<size_t>(f) + (size_t)h * pitch;   // never read past the frame
      for (int v = tid & 31; v < nvec; v += 32) {
        const size_t addr = a0 + (size_t)v * 16;
        uint4 q;
        if (addr + 16 <= row_end && addr >= reinterpret_cast<size_t>(f)) q = __ldg(reinterpret_cast<const uint4*>(addr));
        else {
          uint8_t t8[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            size_t ad = addr + k;
            t8[k] = (ad >= reinterpret_cast<size_t>(f) && ad < row_end) ? *reinterpret_cast<const uint8_t*>(ad) : 0;
          }
          q = *reinterpret_cast<uint4*>(t8);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)v * 16) = q;
      }
    }
  } else {
    // NV12: convert to BGR bytes while staging (BT.601 limited range, cv2's fixed point)
    for (int it = tid; it < rows * ncols; it += NT) {
      const int r = it / ncols, x = it - r * ncols;
#pragma unroll
      for (int c = 0; c < 3; ++c) src[(size_t)r * srcb_cap + x * 3 + c] = (uint8_t)nv12_bgr(f, h, pitch, r0 + r, c0 + x, c);
    }
  }
  __syncthreads();

  // ---- horizontal pass: one item = (source row, output column), all three channels; warps walk rows, lanes columns.
  // Fast path (<= 4 taps): the 12 source bytes of an item come in as four aligned 32-bit words realigned with funnel
  // shifts and the four coefficients as one 16-byte load -- 5 shared-memory loads instead of 16 (the kernel was bound
  // by the issue rate of byte loads, not by HBM); taps past cnt have zero coefficients.
  for (int r = tid >> 5; r < rows; r += NT / 32) {
    int skew = 0;
    if (PIX == 0) skew = (int)((reinterpret_cast<size_t>(f + (size_t)(r0 + r) * pitch) + c0 * 3) & 15);
    const int rowoff = r * srcb_cap + skew - c0 * 3;                 // byte offset of source column 0 of this row in src
    for (int xo = tid & 31; xo < tw; xo += 32) {
      const int lo = sx_lo[xo];
      int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
      if (fastx) {
        const int off = rowoff + lo * 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + (off & ~3));
        const uint32_t sh = (uint32_t)(off & 3) * 8u;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
        const int4 k = *reinterpret_cast<const int4*>(sx_k + xo * 4);
        a0 += (int)(v0 & 255u) * k.x + (int)(v0 >> 24) * k.y + (int)((v1 >> 16) & 255u) * k.z + (int)((v2 >> 8) & 255u) * k.w;
        a1 += (int)((v0 >> 8) & 255u) * k.x + (int)(v1 & 255u) * k.y + (int)(v1 >> 24) * k.z + (int)((v2 >> 16) & 255u) * k.w;
        a2 += (int)((v0 >> 16) & 255u) * k.x + (int)((v1 >> 8) & 255u) * k.y + (int)(v2 & 255u) * k.z + (int)(v2 >> 24) * k.w;
      } else {
        const int cnt = sx_cnt[xo];
        const int* kk = sx_k + xo * ksx;
        const uint8_t* p = src + rowoff + lo * 3;
        for (int j = 0; j < cnt; ++j) {
          const int k = kk[j];
          a0 += (int)p[j * 3] * k; a1 += (int)p[j * 3 + 1] * k; a2 += (int)p[j * 3 + 2] * k;
        }
      }
      tmp[r * TW + xo] = (uint32_t)clip8(a0 >> 22) | ((uint32_t)clip8(a1 >> 22) << 8) | ((uint32_t)clip8(a2 >> 22) << 16);
    }
  }
  __syncthreads();

  // ---- vertical pass + normalise: items = th x TW pixels
  for (int it = tid; it < th * TW; it += NT) {
    const int xo = it & (TW - 1), yo = it / TW;
    if (xo >= tw) continue;
    const int oy = oy0 + yo;
    const int lo = sy_lo[yo] - r0;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    const uint32_t* p = tmp + lo * TW + xo;
    if (fasty) {
      const int4 k = *reinterpret_cast<const int4*>(sy_k + yo * 4);
      const uint32_t t0 = p[0], t1 = p[TW], t2 = p[2 * TW], t3 = p[3 * TW];      // rows past cnt: zero coefficients
      a0 += (int)(t0 & 255u) * k.x + (int)(t1 & 255u) * k.y + (int)(t2 & 255u) * k.z + (int)(t3 & 255u) * k.w;
      a1 += (int)((t0 >> 8) & 255u) * k.x + (int)((t1 >> 8) & 255u) * k.y + (int)((t2 >> 8) & 255u) * k.z +
            (int)((t3 >> 8) & 255u) * k.w;
      a2 += (int)((t0 >> 16) & 255u) * k.x + (int)((t1 >> 16) & 255u) * k.y + (int)((t2 >> 16) & 255u) * k.z +
            (int)((t3 >> 16) & 255u) * k.w;
    } else {
      const int cnt = sy_cnt[yo];
      const int* kk = sy_k + yo * ksy;
      for (int j = 0; j < cnt; ++j) {
        const int k = kk[j];
        const uint32_t t = p[j * TW];
        a0 += (int)(t & 255u) * k; a1 += (int)((t >> 8) & 255u) * k; a2 += (int)((t >> 16) & 255u) * k;
      }
    }
    // source order is B,G,R; the network wants R,G,B (cvtColor at text_detector.py:120).  The table holds, per
    // channel, ToTensor (x/255) followed by Normalize ((x-mean)/std) evaluated with IEEE fp32 ops (see lut kernel).
    const float vr = lut[clip8(a2 >> 22)], vg = lut[256 + clip8(a1 >> 22)], vb = lut[512 + clip8(a0 >> 22)];
    size_t o = (size_t)(lay.offset + blockIdx.z * lay.img_pitch + oy * lay.row_pitch + (long long)(ox0 + xo) * 4);
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(vr, vg, vb, 0.f);
    } else {
      bf16x2 p0 = pack2(vr, vg), p1 = pack2(vb, 0.f);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + o) = u;
    }
  }
}

__global__ void normalize_lut_kernel(float* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 768) return;
  const int c = i >> 8;
  const float x = (float)(i & 255);
  const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
  const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
  lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.0f), mean), sd);
}

}  // namespace

template <typename T>
cudaError_t preprocess_frames(const uint8_t* const* frames_dev, int n, int h, int w, int pitch, int pixfmt,
                              const ResizeTab& tx, const ResizeTab& ty, const float* lut, T* out, OutLayout lay,
                              cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  // source rows / columns one tile can need: T output samples span at most T*scale + ksize source samples
  const double sy = (double)ty.in_size / ty.out_size, sx = (double)tx.in_size / tx.out_size;
  const int rows_cap = (int)(TH * (sy > 1.0 ? sy : 1.0)) + ty.ksize + 2;
  const int cols_cap = (int)(TW * (sx > 1.0 ? sx : 1.0)) + tx.ksize + 2;
  const int srcb_cap = ((cols_cap * 3 + 15 + 15) / 16 + 1) * 16;          // + alignment skew, rounded to 16 bytes
  const int ksx = tx.maxcnt <= 4 ? 4 : tx.ksize, ksy = ty.maxcnt <= 4 ? 4 : ty.ksize;
  size_t smem = sizeof(int) * (size_t)(2 * TW + TW * ksx + 2 * TH + TH * ksy);
  smem = ((smem + 15) & ~(size_t)15) + 768 * sizeof(float);
  smem += (size_t)rows_cap * srcb_cap + (size_t)(rows_cap + 4) * TW * 4 + 16;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  dim3 grid((tx.out_size + TW - 1) / TW, (ty.out_size + TH - 1) / TH, n);
  cudaError_t e;
  if (pixfmt == 0) {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(preprocess_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    preprocess_kernel<T, 0><<<grid, NT, smem, s>>>(frames_dev, h, w, pitch, tx, ty, rows_cap, srcb_cap, lut, out, lay);
  } else {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(preprocess_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    preprocess_kernel<T, 1><<<grid, NT, smem, s>>>(frames_dev, h, w, pitch, tx, ty, rows_cap, srcb_cap, lut, out, lay);
  }
  if (lc) lc->n++;
  return cudaGetLastError();
}

template cudaError_t preprocess_frames<float>(const uint8_t* const*, int, int, int, int, int, const ResizeTab&,
                                              const ResizeTab&, const float*, float*, OutLayout, cudaStream_t, LaunchCounter*);
template cudaError_t preprocess_frames<bf16>(const uint8_t* const*, int, int, int, int, int, const ResizeTab&,
                                             const ResizeTab&, const float*, bf16*, OutLayout, cudaStream_t, LaunchCounter*);

cudaError_t build_normalize_lut(float* lut_dev, cudaStream_t s) {
  normalize_lut_kernel<<<3, 256, 0, s>>>(lut_dev);
  return cudaGetLastError();
}

}  // namespace vtd
