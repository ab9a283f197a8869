// Stage 3: fused DB head tail.
//
// DBHead (text_detector.py:58-86) is, per branch, Conv3x3(256->64)+BN+ReLU -> ConvT2x2s2(64->64)+BN+ReLU
// -> ConvT2x2s2(64->1) -> Sigmoid.  A k=2,s=2 transposed conv has no overlap between outputs, so after the
// 3x3 conv everything is local to one stride-4 pixel: 64 -> (2x2 positions x 64) -> ReLU -> (4x4 x 1) ->
// sigmoid.  The 3x3 convs of both branches run as ONE 256->128 implicit GEMM (conv kernels); this kernel
// takes that [N,H/4,W/4,128] feature map and emits, in one HBM pass, probability (fp32), threshold (fp32)
// and the binary mask `prob > thr` (u8; the strict '>' of text_detector.py:144).  The four intermediate
// maps of the reference never exist in memory.
//
// CTA = 256 threads, persistent over tiles of 32 consecutive stride-4 pixels of one row, one branch per
// blockIdx.y.  W1 (64x256 fp32, BN folded) stays in shared memory for the CTA's lifetime.  Outputs are
// staged in shared memory and written as full-row float4 vectors (4 output rows x 128 floats per tile).
// Algorithmic bytes per frame: read H4*W4*128*sizeof(T), write 2*Hd*Wd*4 + Hd*Wd.
#include "common.cuh"

namespace vtd {
namespace {

constexpr int TP = 32;       // stride-4 pixels per tile
constexpr int NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) db_head_tail_kernel(const T* __restrict__ feat, HeadTailWeights hw, int N,
                                                          int H4, int W4, const float* __restrict__ logit_bias,
                                                          float thr, float* __restrict__ prob,
                                                          float* __restrict__ thresh, uint8_t* __restrict__ mask) {
  extern __shared__ float smem[];
  float* W1s = smem;                    // [64 ci][256 co']
  float* Fs = W1s + 64 * 256;           // [64 ci][TP+4]
  float* Os = Fs + 64 * (TP + 4);       // [TP][16]
  const int head = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;

  // W1 -> smem, transposed to [ci][co']
  {
    const float* w = hw.w1 + (size_t)head * 256 * 64 + (size_t)tid * 64;
#pragma unroll 4
    for (int c4 = 0; c4 < 16; ++c4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(w) + c4);
      W1s[(c4 * 4 + 0) * 256 + tid] = v.x; W1s[(c4 * 4 + 1) * 256 + tid] = v.y;
      W1s[(c4 * 4 + 2) * 256 + tid] = v.z; W1s[(c4 * 4 + 3) * 256 + tid] = v.w;
    }
  }
  float b1[8], w2[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int co = tx * 8 + j;                 // (dy,dx,c): group g = co/64, c = co%64
    b1[j] = __ldg(hw.b1 + head * 256 + co);
    const float* q = hw.w2 + ((size_t)head * 64 + (co & 63)) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) w2[j][k] = __ldg(q + k);
  }
  const float b2 = __ldg(hw.b2 + head);
  float* __restrict__ outp = head == 0 ? prob : thresh;
  const int Wd = 4 * W4, Hd = 4 * H4;
  const int tiles_per_row = (W4 + TP - 1) / TP;
  const long long ntiles = (long long)N * H4 * tiles_per_row;
  __syncthreads();

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int txi = (int)(tile % tiles_per_row);
    const long long row = tile / tiles_per_row;       // n*H4 + y
    const int y = (int)(row % H4), n = (int)(row / H4);
    const int x0 = txi * TP;
    const int tp = min(TP, W4 - x0);
    // feature tile -> smem [ci][px]
    const T* f = feat + ((size_t)row * W4 + x0) * 128 + head * 64;
    for (int it = tid; it < TP * 64; it += NT) {
      int ci = it & 63, px = it >> 6;
      Fs[ci * (TP + 4) + px] = px < tp ? to_f(f[(size_t)px * 128 + ci]) : 0.f;
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int ci = 0; ci < 64; ++ci) {
      float4 a = *reinterpret_cast<const float4*>(&Fs[ci * (TP + 4) + ty * 4]);
      float4 w0 = *reinterpret_cast<const float4*>(&W1s[ci * 256 + tx * 8]);
      float4 w1v = *reinterpret_cast<const float4*>(&W1s[ci * 256 + tx * 8 + 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    // second transposed conv: partial over this thread's 8 channels, reduce over the 8 lanes of the group
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float hval = fmaxf(acc[i][j] + b1[j], 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = fmaf(hval, w2[j][k], p[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        p[k] += __shfl_xor_sync(0xffffffffu, p[k], 1);
        p[k] += __shfl_xor_sync(0xffffffffu, p[k], 2);
        p[k] += __shfl_xor_sync(0xffffffffu, p[k], 4);
      }
      if ((tx & 7) == 0) {
        int g = tx >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) Os[(ty * 4 + i) * 16 + g * 4 + k] = p[k] + b2;
      }
    }
    __syncthreads();
    // write 4 output rows x (tp*4) floats, one float4 (= one stride-4 pixel's row of 4) per item
    for (int it = tid; it < 4 * tp; it += NT) {
      int px = it % tp, r = it / tp;
      int dy = r >> 1, dy2 = r & 1;
      const float* o = Os + px * 16;
      float v[4];
      v[0] = o[(dy * 2 + 0) * 4 + dy2 * 2 + 0];
      v[1] = o[(dy * 2 + 0) * 4 + dy2 * 2 + 1];
      v[2] = o[(dy * 2 + 1) * 4 + dy2 * 2 + 0];
      v[3] = o[(dy * 2 + 1) * 4 + dy2 * 2 + 1];
      size_t oidx = ((size_t)n * Hd + 4 * y + r) * Wd + 4 * (x0 + px);
      if (head == 0 && logit_bias) {
        float4 lb = __ldg(reinterpret_cast<const float4*>(logit_bias + oidx));
        v[0] += lb.x; v[1] += lb.y; v[2] += lb.z; v[3] += lb.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = 1.0f / (1.0f + expf(-v[k]));
      *reinterpret_cast<float4*>(outp + oidx) = make_float4(v[0], v[1], v[2], v[3]);
      if (head == 0) {
        uint32_t m = (v[0] > thr ? 1u : 0u) | (v[1] > thr ? 0x100u : 0u) | (v[2] > thr ? 0x10000u : 0u) |
                     (v[3] > thr ? 0x1000000u : 0u);
        *reinterpret_cast<uint32_t*>(mask + oidx) = m;
      }
    }
    __syncthreads();
  }
}

}  // namespace

template <typename T>
cudaError_t db_head_tail(const T* feat, const HeadTailWeights& hw, int N, int H4, int W4, const float* logit_bias,
                         float thr, float* prob, float* thresh, uint8_t* mask, cudaStream_t s, LaunchCounter* lc) {
  if (N <= 0) return cudaSuccess;
  size_t smem = sizeof(float) * (64 * 256 + 64 * (TP + 4) + TP * 16);
  static PerDeviceFlag attr_done;                 // one flag per instantiation (T)
  {
    cudaError_t e = once_per_device(attr_done, [smem] {
      return cudaFuncSetAttribute(db_head_tail_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
    if (e != cudaSuccess) return e;
  }
  long long ntiles = (long long)N * H4 * ((W4 + TP - 1) / TP);
  int gx = (int)(ntiles < 148 * 2 ? ntiles : 148 * 2);
  db_head_tail_kernel<T><<<dim3(gx, 2), NT, smem, s>>>(feat, hw, N, H4, W4, logit_bias, thr, prob, thresh, mask);
  if (lc) lc->n++;
  return cudaGetLastError();
}

template cudaError_t db_head_tail<float>(const float*, const HeadTailWeights&, int, int, int, const float*, float,
                                         float*, float*, uint8_t*, cudaStream_t, LaunchCounter*);
template cudaError_t db_head_tail<bf16>(const bf16*, const HeadTailWeights&, int, int, int, const float*, float,
                                        float*, float*, uint8_t*, cudaStream_t, LaunchCounter*);

}  // namespace vtd
