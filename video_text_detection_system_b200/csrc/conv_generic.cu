// CUDA-core implicit-GEMM convolution over NHWC activations (any kernel size / stride / channel count).
//
// Role: (1) the fp32 parity tier (VTD_FP32): every conv of DBNet (text_detector.py:12-86) and CRNN
// (text_recognizer.py:16-27) computed with fp32 FFMA and fp32 accumulation, <=1e-3 of the reference's
// eager fp32 path; (2) in the bf16 tier, the layers the tcgen05 kernel does not take (Cin not a
// multiple of 64: the two 3-channel stems).
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin with k = (r*KW+s)*Cin+ci.
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, register-prefetched global loads.
#include "common.cuh"

namespace vtd {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <typename T> struct Vec8;   // 8 consecutive channels
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Vec8<bf16> {
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const bf16x2* h = reinterpret_cast<const bf16x2*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = unpack2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float* v);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* v) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float* v) {
  uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  const bf16x2* h = reinterpret_cast<const bf16x2*>(&a);
  float2 f0 = unpack2(h[0]), f1 = unpack2(h[1]);
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}

template <typename T> __device__ __forceinline__ void store4(T* p, const float* v);
template <> __device__ __forceinline__ void store4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, const float* v) {
  bf16x2 a = pack2(v[0], v[1]), b = pack2(v[2], v[3]);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// FAST: Cin % 16 == 0 (a 16-wide K chunk never straddles a tap; 16B-aligned vector loads).
template <typename T, bool FAST>
__global__ void __launch_bounds__(NT) conv_generic_kernel(ConvDesc d) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const T* __restrict__ in = reinterpret_cast<const T*>(d.in);
  const T* __restrict__ wt = reinterpret_cast<const T*>(d.w);
  const int tid = threadIdx.x;
  const int K = d.KH * d.KW * d.Cin;
  const long long M = (long long)d.N * d.Ho * d.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader: row = tid/2, 8 k-elements starting at (tid%2)*8
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const long long am = m0 + a_row;
  const bool a_valid = am < M;
  int an = 0, aoy = 0, aox = 0;
  if (a_valid) {
    an = (int)(am / ((long long)d.Ho * d.Wo));
    int rem = (int)(am - (long long)an * d.Ho * d.Wo);
    aoy = rem / d.Wo; aox = rem - aoy * d.Wo;
  }
  const int iy0 = aoy * d.stride - d.pad, ix0 = aox * d.stride - d.pad;
  // ---- B loader: row = tid/4 (cout), 4 k-elements starting at (tid%4)*4
  const int b_row = tid >> 2, b_k = (tid & 3) * 4;
  const int bco = n0 + b_row;
  const bool b_valid = bco < d.Cout;

  float ra[8], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ra[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) rb[j] = 0.f;
    if (FAST) {
      if (a_valid) {
        int tap = k0 / d.Cin, ci0 = k0 - tap * d.Cin;
        int r = tap / d.KW, s = tap - r * d.KW;
        int iy = iy0 + r, ix = ix0 + s;
        if ((unsigned)iy < (unsigned)d.H && (unsigned)ix < (unsigned)d.W) {
          Vec8<T> v; v.load(in + (((long long)an * d.H + iy) * d.W + ix) * d.Cin + ci0 + a_k);
#pragma unroll
          for (int j = 0; j < 8; ++j) ra[j] = v.v[j];
        }
      }
      if (b_valid) load4<T>(wt + (long long)bco * K + k0 + b_k, rb);
    } else {
      if (a_valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int k = k0 + a_k + j;
          if (k < K) {
            int tap = k / d.Cin, ci = k - tap * d.Cin;
            int r = tap / d.KW, s = tap - r * d.KW;
            int iy = iy0 + r, ix = ix0 + s;
            if ((unsigned)iy < (unsigned)d.H && (unsigned)ix < (unsigned)d.W)
              ra[j] = to_f(in[(((long long)an * d.H + iy) * d.W + ix) * d.Cin + ci]);
          }
        }
      }
      if (b_valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = k0 + b_k + j;
          if (k < K) rb[j] = to_f(wt[(long long)bco * K + k]);
        }
      }
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  for (int kb = 0; kb < nk; ++kb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[a_k + j][a_row] = ra[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[b_k + j][b_row] = rb[j];
    __syncthreads();
    if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias + residual + ReLU, NHWC or depth-to-space store
  const int co0 = n0 + tx * 4;
  if (co0 >= d.Cout) return;
  const bool vec_ok = (co0 + 3 < d.Cout) && (d.Cout % 4 == 0) &&
                      (d.out_mode == OUT_NHWC || (d.Cout / 4) % 4 == 0);
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias[j] = (co0 + j < d.Cout && d.bias) ? __ldg(d.bias + co0 + j) : 0.f;
  const T* __restrict__ res = reinterpret_cast<const T*>(d.res);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + ty * 8 + i;
    if (m >= M) break;
    int n = (int)(m / ((long long)d.Ho * d.Wo));
    int rem = (int)(m - (long long)n * d.Ho * d.Wo);
    int oy = rem / d.Wo, ox = rem - oy * d.Wo;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (d.res_mode != RES_NONE) {
      long long ridx = (d.res_mode == RES_SAME)
                           ? (((long long)n * d.Ho + oy) * d.Wo + ox) * d.Cout + co0
                           : (((long long)n * (d.Ho / 2) + (oy >> 1)) * (d.Wo / 2) + (ox >> 1)) * d.Cout + co0;
      if (vec_ok) { float r[4]; load4<T>(res + ridx, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += r[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (co0 + j < d.Cout) v[j] += to_f(res[ridx + j]);
      }
    }
    if (d.relu == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (d.relu == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
    }
    if (d.out_mode == OUT_NHWC) {
      long long oidx = (((long long)n * d.Ho + oy) * d.Wo + ox) * d.Cout + co0;
      if (d.out_f32) {
        float* o = reinterpret_cast<float*>(d.out);
        if (vec_ok) store4<float>(o + oidx, v);
        else
#pragma unroll
          for (int j = 0; j < 4; ++j) if (co0 + j < d.Cout) o[oidx + j] = v[j];
      } else {
        T* o = reinterpret_cast<T*>(d.out);
        if (vec_ok) store4<T>(o + oidx, v);
        else
#pragma unroll
          for (int j = 0; j < 4; ++j) if (co0 + j < d.Cout) o[oidx + j] = from_f<T>(v[j]);
      }
    } else {  // OUT_D2S
      const int C2 = d.Cout / 4;
      T* o = reinterpret_cast<T*>(d.out);
      if (vec_ok) {
        int q = co0 / C2, c = co0 - q * C2;
        long long oidx = (((long long)n * 2 * d.Ho + 2 * oy + (q >> 1)) * (2 * d.Wo) + 2 * ox + (q & 1)) * C2 + c;
        store4<T>(o + oidx, v);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int co = co0 + j;
          if (co < d.Cout) {
            int q = co / C2, c = co - q * C2;
            long long oidx = (((long long)n * 2 * d.Ho + 2 * oy + (q >> 1)) * (2 * d.Wo) + 2 * ox + (q & 1)) * C2 + c;
            o[oidx] = from_f<T>(v[j]);
          }
        }
      }
    }
  }
}

}  // namespace

template <typename T>
cudaError_t conv_generic(const ConvDesc& d, cudaStream_t s, LaunchCounter* lc) {
  const long long M = (long long)d.N * d.Ho * d.Wo;
  if (M <= 0) return cudaSuccess;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d.Cout + BN - 1) / BN));
  const bool fast = (d.Cin % 16 == 0);
  if (fast) conv_generic_kernel<T, true><<<grid, NT, 0, s>>>(d);
  else conv_generic_kernel<T, false><<<grid, NT, 0, s>>>(d);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template cudaError_t conv_generic<float>(const ConvDesc&, cudaStream_t, LaunchCounter*);
template cudaError_t conv_generic<bf16>(const ConvDesc&, cudaStream_t, LaunchCounter*);

}  // namespace vtd
