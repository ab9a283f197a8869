// Bidirectional LSTM layer of the CRNN recogniser (text_recognizer.py:26: nn.LSTM(512, 256, 2 layers,
// bidirectional, batch_first)).  PyTorch gate order i,f,g,o; c' = f*c + i*g; h' = o*tanh(c').
//
// The input projection W_ih x + b_ih + b_hh of all timesteps and both directions is one implicit GEMM
// (the conv kernels, 1x1) that leaves xproj [B,T,2,4H] fp32.  This file is the recurrence: per timestep
// one launch computes, for both directions, gates = xproj[t] + h_{t-1} W_hh^T as a tiled GEMM with the cell
// update fused into its epilogue (a CTA owns 32 sequences x 32 hidden units x all 4 gates, so i,f,g,o of a
// unit meet in one thread).  h is ping-ponged between two buffers; W_hh (2 x 1024 x 256) stays L2-resident.
#include "common.cuh"

namespace vtd {
namespace {

constexpr int TB = 32, TJ = 32, TK = 32, NT = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T, typename WT>
__global__ void __launch_bounds__(NT) lstm_step_kernel(const float* __restrict__ xproj, const WT* __restrict__ whh,
                                                       const float* __restrict__ h_prev, float* __restrict__ h_next,
                                                       float* __restrict__ cbuf, T* __restrict__ out, int B, int Tn,
                                                       int H, int step) {
  __shared__ float Hs[TK][TB + 1];
  __shared__ float Ws[4][TK][TJ + 1];
  const int dir = blockIdx.z;
  const int b0 = blockIdx.x * TB, j0 = blockIdx.y * TJ;
  const int t = dir == 0 ? step : Tn - 1 - step;
  const int tid = threadIdx.x;
  const int tb = tid >> 3;            // sequence within the tile
  const int tj = (tid & 7) * 4;       // first of 4 hidden units
  const float* hp = h_prev + (size_t)dir * B * H;
  const WT* w = whh + (size_t)dir * 4 * H * H;
  float acc[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[g][j] = 0.f;

  if (step > 0) {
    for (int k0 = 0; k0 < H; k0 += TK) {
      // h tile: 32 sequences x 32 k
      for (int it = tid; it < TB * TK; it += NT) {
        int k = it & (TK - 1), b = it / TK;
        Hs[k][b] = (b0 + b < B) ? hp[(size_t)(b0 + b) * H + k0 + k] : 0.f;
      }
      // W tile: 4 gates x 32 units x 32 k   (W_hh row = gate*H + unit, col = k)
      for (int it = tid; it < 4 * TJ * TK; it += NT) {
        int k = it & (TK - 1), j = (it / TK) & (TJ - 1), g = it / (TK * TJ);
        Ws[g][k][j] = to_f(w[((size_t)g * H + j0 + j) * H + k0 + k]);
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < TK; ++k) {
        float hv = Hs[k][tb];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[g][j] = fmaf(hv, Ws[g][k][tj + j], acc[g][j]);
      }
      __syncthreads();
    }
  }
  const int b = b0 + tb;
  if (b >= B) return;
  const float* xp = xproj + (((size_t)b * Tn + t) * 2 + dir) * 4 * H;
  float* c = cbuf + ((size_t)dir * B + b) * H;
  float* hn = h_next + ((size_t)dir * B + b) * H;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int u = j0 + tj + j;
    float gi = acc[0][j] + xp[u];
    float gf = acc[1][j] + xp[H + u];
    float gg = acc[2][j] + xp[2 * H + u];
    float go = acc[3][j] + xp[3 * H + u];
    float cp = step > 0 ? c[u] : 0.f;
    float cn = sigmoidf_(gf) * cp + sigmoidf_(gi) * tanhf(gg);
    float hv = sigmoidf_(go) * tanhf(cn);
    c[u] = cn;
    hn[u] = hv;
    out[((size_t)b * Tn + t) * 2 * H + (size_t)dir * H + u] = from_f<T>(hv);
  }
}

}  // namespace

template <typename T, typename WT>
cudaError_t bilstm_layer(const float* xproj, const WT* whh, T* out, float* hbuf, float* cbuf, int B, int Tn, int H,
                         cudaStream_t s, LaunchCounter* lc) {
  if (B <= 0 || Tn <= 0) return cudaSuccess;
  if (H % TJ != 0) return cudaErrorInvalidValue;
  dim3 grid((B + TB - 1) / TB, H / TJ, 2);
  const size_t hsz = (size_t)2 * B * H;
  for (int step = 0; step < Tn; ++step) {
    const float* hp = hbuf + (size_t)(step & 1) * hsz;
    float* hn = hbuf + (size_t)((step + 1) & 1) * hsz;
    lstm_step_kernel<T, WT><<<grid, NT, 0, s>>>(xproj, whh, hp, hn, cbuf, out, B, Tn, H, step);
  }
  if (lc) lc->n += Tn;
  return cudaGetLastError();
}

template cudaError_t bilstm_layer<float, float>(const float*, const float*, float*, float*, float*, int, int, int,
                                                cudaStream_t, LaunchCounter*);
template cudaError_t bilstm_layer<bf16, bf16>(const float*, const bf16*, bf16*, float*, float*, int, int, int,
                                              cudaStream_t, LaunchCounter*);

}  // namespace vtd
