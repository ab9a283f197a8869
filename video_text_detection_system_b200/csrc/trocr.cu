// Transformer recogniser (TrOCR branch of the reference: app/ml/models/text_recognizer.py:39-69): the kernels that are not
// GEMMs.  The reference calls HuggingFace's VisionEncoderDecoderModel (ViT-B/16 @384 encoder, 12-layer TrOCR decoder,
// greedy generate, max_length 50); every Linear / patch-embedding projection of that model runs on the tcgen05 implicit-GEMM
// kernels of conv_tcgen05.cu (1x1 "convolutions" over token maps, bias / residual / GELU fused in their epilogues).  This
// file holds the rest, all 16-bit storage with fp32 arithmetic:
//   trocr_resize_patches : per-crop Pillow-exact antialiased bilinear resize to 384x384 (TrOCRProcessor: resample=2),
//                          BGR->RGB, /255, (x-0.5)/0.5, written straight in patch order (the A operand of the patch GEMM)
//   vit_assemble         : [CLS] + patch embeddings + position embeddings
//   layernorm_rows       : nn.LayerNorm over the channel dimension (fp32 statistics)
//   attention_enc        : bidirectional multi-head attention of the encoder, flash-style (online softmax, K/V streamed
//                          through shared memory), QK^T and PV on mma.sync m16n8k16 tensor-core instructions
//   attention_decode     : one query per (crop, head) against a K/V cache (decoder self- and cross-attention)
//   trocr_embed          : token embedding + learned position embedding (offset 2)
//   kv_append            : the step's K,V into the self-attention cache
//   argmax_rows          : greedy token choice over the vocabulary (lowest index on ties, torch.argmax), EOS/pad handling
#include "common.cuh"

namespace vtd {
namespace {

#ifdef VTD_BF16_STORAGE
#define VTD_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32"
#else
#define VTD_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32"
#endif

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(VTD_MMA_16816 " {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  bf16x2 h = pack2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch for the decode loop's chain of small kernels: every kernel lets its successor start at once
// (pdl_go) and waits for its predecessor's results only where it first needs them (pdl_wait) -- block scheduling, the prologue
// and, in the skinny GEMM, the whole weight prefetch overlap the predecessor's tail.  A kernel launched without the attribute
// (or after a non-kernel stream operation) finds both instructions to be no-ops.
__device__ __forceinline__ void pdl_go() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- LayerNorm: one warp per row ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y, long long rows, int C,
                                                        float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16x2* xr = reinterpret_cast<const bf16x2*>(x + row * C);
  const int C2 = C >> 1;
  float s = 0.f;
  for (int i = lane; i < C2; i += 32) { const float2 v = unpack2(xr[i]); s += v.x + v.y; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int i = lane; i < C2; i += 32) { const float2 v = unpack2(xr[i]); const float a = v.x - mean, b = v.y - mean; q += a * a + b * b; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  const float rstd = rsqrtf(q / (float)C + eps);
  bf16x2* yr = reinterpret_cast<bf16x2*>(y + row * C);
  for (int i = lane; i < C2; i += 32) {
    const float2 v = unpack2(xr[i]);
    const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * i), b = *reinterpret_cast<const float2*>(beta + 2 * i);
    yr[i] = pack2((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y);
  }
}

// LayerNorm of a handful of rows (the decode loop: one row per crop): one CTA per row, the row held in registers, so the
// two reductions cost two block-wide shuffles instead of three strided passes of one warp.
__global__ void __launch_bounds__(256) layernorm_row_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y, int C, float eps) {
  __shared__ float red[2][8];
  pdl_go();
  pdl_wait();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const size_t row = blockIdx.x;
  const bool on = t * 8 < C;
  float v[8];
  float s = 0.f;
  if (on) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + t * 8);
    const bf16x2* h = reinterpret_cast<const bf16x2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = unpack2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; s += f.x + f.y; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[0][w];
  const float mean = s / (float)C;
  float q = 0.f;
  if (on) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float a = v[e] - mean; q += a * a; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  q = 0.f;
  for (int w = 0; w < nw; ++w) q += red[1][w];
  const float rstd = rsqrtf(q / (float)C + eps);
  if (on) {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + t * 8), g1 = *reinterpret_cast<const float4*>(gamma + t * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + t * 8), b1 = *reinterpret_cast<const float4*>(beta + t * 8 + 4);
    uint4 o;
    o.x = pack16((v[0] - mean) * rstd * g0.x + b0.x, (v[1] - mean) * rstd * g0.y + b0.y);
    o.y = pack16((v[2] - mean) * rstd * g0.z + b0.z, (v[3] - mean) * rstd * g0.w + b0.w);
    o.z = pack16((v[4] - mean) * rstd * g1.x + b1.x, (v[5] - mean) * rstd * g1.y + b1.y);
    o.w = pack16((v[6] - mean) * rstd * g1.z + b1.z, (v[7] - mean) * rstd * g1.w + b1.w);
    *reinterpret_cast<uint4*>(y + row * C + t * 8) = o;
  }
}

// ---- skinny GEMM: the Linears of the DECODE loop ------------------------------------------------------------------------
// Y[m][n] = act(sum_k X[m][k] W[n][k] + bias[n]) (+ res[m][n]) for m < M <= 128 rows (one row per crop).  On the 128-row tcgen05
// tiles such a GEMM is N/256 CTAs streaming 0.5-2 MB of weights each (19 us per launch measured); here the N columns are spread
// over every SM -- a CTA owns 8*WN columns, its 8 warps are WN column tiles x (8/WN) slices of K -- and the weights cross HBM
// once at full width.  mma.sync m16n8k16 with a permuted K order: within a 32-wide K block thread t of a quad loads the 8
// consecutive elements 8t..8t+7 of its A rows and of its B column as ONE 16-byte load each and feeds elements 0..3 to the first
// MMA, 4..7 to the second; A and B use the same permutation, so the sum over K is unchanged.  Partial sums of the K slices meet
// in shared memory, where bias, residual and activation are applied.
template <int MT>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const bf16* __restrict__ X, int ldx, const bf16* __restrict__ W,
                                                          const float* __restrict__ bias, const bf16* __restrict__ res, int ldres,
                                                          void* __restrict__ out, int ldo, int out_f32, int M, int N, int K, int act,
                                                          int WN) {
  __shared__ float red[MT * 16 * 64];                    // [K slice][row][column of the CTA]
  extern __shared__ __align__(16) uint8_t wsm[];         // the CTA's 8*WN weight rows, pitch 2K + 64 bytes (conflict-free 16-byte reads)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int WK = 8 / WN, wn = warp % WN, wk = warp / WN;
  const int cols = 8 * WN;
  const int ks = K / WK, k0 = wk * ks;
  // the whole weight slice is requested at once (cp.async): 16-130 KB in flight per SM is what keeps HBM busy when every CTA
  // has only microseconds of work
  const int pitch = 2 * K + 64, chunks = K / 8;
  pdl_go();
  for (int i = threadIdx.x; i < cols * chunks; i += 256) {
    const int r = i / chunks, ch = i % chunks;
    const int col = blockIdx.x * cols + r;
    const bf16* src = W + (size_t)(col < N ? col : N - 1) * K + ch * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(wsm + (size_t)r * pitch + ch * 16)), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  pdl_wait();                                            // the weights do not depend on the previous kernel; X, res and out do
  float acc[MT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  const uint8_t* wrow = wsm + (size_t)(wn * 8 + g) * pitch + (size_t)(k0 + t * 8) * 2;
  const bf16* xrow[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int r0 = i * 16 + g, r1 = r0 + 8;
    xrow[i][0] = X + (size_t)(r0 < M ? r0 : M - 1) * ldx + k0 + t * 8;
    xrow[i][1] = X + (size_t)(r1 < M ? r1 : M - 1) * ldx + k0 + t * 8;
  }
  // two 32-wide K blocks of activations are always in flight, and the first two are requested BEFORE the weights are waited
  // for: the kernel's time is its chain of memory round trips, not its bytes
  uint4 xa[2][MT], xb[2][MT];
  auto load_x = [&](int st, int kb) {
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      xa[st][i] = *reinterpret_cast<const uint4*>(xrow[i][0] + kb);
      xb[st][i] = *reinterpret_cast<const uint4*>(xrow[i][1] + kb);
    }
  };
  auto mma_block = [&](int st, int kb) {
    const uint4 wv = *reinterpret_cast<const uint4*>(wrow + (size_t)kb * 2);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const uint32_t a1[4] = {xa[st][i].x, xb[st][i].x, xa[st][i].y, xb[st][i].y};
      const uint32_t a2[4] = {xa[st][i].z, xb[st][i].z, xa[st][i].w, xb[st][i].w};
      mma16816(acc[i], a1, wv.x, wv.y);
      mma16816(acc[i], a2, wv.z, wv.w);
    }
  };
  load_x(0, 0);
  if (ks > 32) load_x(1, 32);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int kb = 0; kb < ks; kb += 64) {
    mma_block(0, kb);
    if (kb + 64 < ks) load_x(0, kb + 64);
    if (kb + 32 < ks) {
      mma_block(1, kb + 32);
      if (kb + 96 < ks) load_x(1, kb + 96);
    }
  }
  // C fragment: (row g, cols 2t, 2t+1), (row g + 8, cols 2t, 2t+1)
  float* mine = red + (size_t)wk * (MT * 16) * cols;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int r0 = i * 16 + g, c = wn * 8 + 2 * t;
    mine[r0 * cols + c] = acc[i][0]; mine[r0 * cols + c + 1] = acc[i][1];
    mine[(r0 + 8) * cols + c] = acc[i][2]; mine[(r0 + 8) * cols + c + 1] = acc[i][3];
  }
  __syncthreads();
  const int nb = blockIdx.x * cols;
  for (int o = threadIdx.x; o < MT * 16 * cols; o += 256) {
    const int m = o / cols, c = o % cols, col = nb + c;
    if (m >= M || col >= N) continue;
    float v = 0.f;
    for (int w = 0; w < WK; ++w) v += red[(size_t)w * (MT * 16) * cols + o];
    if (bias) v += bias[col];
    if (res) v += to_f(res[(size_t)m * ldres + col]);
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (out_f32) reinterpret_cast<float*>(out)[(size_t)m * ldo + col] = v;
    else reinterpret_cast<bf16*>(out)[(size_t)m * ldo + col] = f32_to_16(v);
  }
}

// ---- encoder attention -------------------------------------------------------------------------------------------------
// grid (ceil(S/64), heads, n); 4 warps, 16 query rows each; head dimension 64.
constexpr int AT_LD = 72;          // shared-memory row pitch in elements (64 + 8: conflict-free ldmatrix)
__global__ void __launch_bounds__(128) attention_enc_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int S, int heads,
                                                            float scale_log2e) {
  __shared__ __align__(16) bf16 Qs[64 * AT_LD];
  __shared__ __align__(16) bf16 Ks[64 * AT_LD];
  __shared__ __align__(16) bf16 Vs[64 * AT_LD];
  const int D = heads * 64, ld = 3 * D;
  const int n = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + (size_t)n * S * ld + hd * 64;
  // stage the 64 query rows (rows beyond S: zeros)
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < S) v = *reinterpret_cast<const uint4*>(base + (size_t)(q0 + r) * ld + c * 8);
    *reinterpret_cast<uint4*>(Qs + r * AT_LD + c * 8) = v;
  }
  __syncthreads();
  uint32_t qf[4][4];                                   // A fragments of this warp's 16 rows, 4 k-steps over d
  {
    const int r = warp * 16 + (lane & 15), cofs = (lane >> 4) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = smem_addr(Qs + r * AT_LD + k * 16 + cofs);
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(qf[k][0]), "=r"(qf[k][1]), "=r"(qf[k][2]), "=r"(qf[k][3]) : "r"(a));
    }
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;    // running max / sum of rows lane/4 and lane/4 + 8
  for (int k0 = 0; k0 < S; k0 += 64) {
    __syncthreads();                                   // the previous tile has been consumed
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
      if (k0 + r < S) {
        const bf16* p = base + (size_t)(k0 + r) * ld + c * 8;
        kv = *reinterpret_cast<const uint4*>(p + D);
        vv = *reinterpret_cast<const uint4*>(p + 2 * D);
      }
      *reinterpret_cast<uint4*>(Ks + r * AT_LD + c * 8) = kv;
      *reinterpret_cast<uint4*>(Vs + r * AT_LD + c * 8) = vv;
    }
    __syncthreads();
    // scores of 16 rows x 64 keys
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t b0, b1;
        const uint32_t a = smem_addr(Ks + (j * 8 + (lane & 7)) * AT_LD + k * 16 + ((lane >> 3) & 1) * 8);
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(a));
        mma16816(sc[j], qf[k], b0, b1);
      }
    }
    // mask keys beyond S, scale, online softmax
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = k0 + j * 8 + 2 * (lane & 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = key + (e & 1) < S;
        sc[j][e] = ok ? sc[j][e] * scale_log2e : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);      // m = -inf on the first tile: exp2f(-inf) = 0
    m0 = mx0; m1 = mx1;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = exp2f(sc[j][0] - mx0); sc[j][1] = exp2f(sc[j][1] - mx0);
      sc[j][2] = exp2f(sc[j][2] - mx1); sc[j][3] = exp2f(sc[j][3] - mx1);
      s0 += sc[j][0] + sc[j][1]; s1 += sc[j][2] + sc[j][3];
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
    }
    l0 = l0 * c0 + s0; l1 = l1 * c1 + s1;
    // O += P V : k-steps of 16 keys, n-tiles of 8 head dimensions
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack16(sc[2 * kk][0], sc[2 * kk][1]); pa[1] = pack16(sc[2 * kk][2], sc[2 * kk][3]);
      pa[2] = pack16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]); pa[3] = pack16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t b0, b1;
        const uint32_t a = smem_addr(Vs + (kk * 16 + (lane & 15)) * AT_LD + j * 8);
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(a));
        mma16816(o[j], pa, b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* ob = out + (size_t)n * S * D + hd * 64 + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < S) *reinterpret_cast<bf16x2*>(ob + (size_t)r0 * D + j * 8) = pack2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < S) *reinterpret_cast<bf16x2*>(ob + (size_t)r1 * D + j * 8) = pack2(o[j][2] * i1, o[j][3] * i1);
  }
}

// ---- decode attention: one query per (crop, head) against L cached keys/values -------------------------------------------
// q: [n][ldq] (head hd at hd*64); k, v: [n][Lcap][ldkv] (head hd at hd*64 from each base pointer); out: [n][heads*64]
constexpr int AD_MAXL = 640;
// Eight lanes share a key / value row: each reads 16 bytes (8 of the 64 head dimensions), so a warp instruction covers four
// whole 128-byte rows, 16 rows per CTA pass, all loads of a pass independent (the rows of one head are 2*D*2 bytes apart in the
// [token][K | V] buffers; one thread per row with eight dependent 16-byte loads ran at 2 TB/s).
constexpr int AD_THREADS = 256, AD_GROUPS = AD_THREADS / 8, AD_WARPS = AD_THREADS / 32;
__global__ void __launch_bounds__(AD_THREADS) attention_decode_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k,
                                                               const bf16* __restrict__ v, int ldkv, long long crop_stride,
                                                               int head_stride, int L, float scale,
                                                               bf16* __restrict__ out, int D, const int* __restrict__ tdev) {
  __shared__ float sc[AD_MAXL];
  pdl_go();
  pdl_wait();
  if (tdev) L = *tdev + 1;                               // decode position on the device (graph replay): keys 0..t
  __shared__ float red[AD_WARPS];
  __shared__ float part[AD_GROUPS][64];
  const int hd = blockIdx.x, n = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int sub = t & 7, grp = t >> 3;                  // 8 dimensions sub*8.., row group 0..AD_GROUPS-1
  float qv[8];
  {
    const uint4 u = *reinterpret_cast<const uint4*>(q + (size_t)n * ldq + hd * 64 + sub * 8);
    const bf16x2* h = reinterpret_cast<const bf16x2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = unpack2(h[e]); qv[2 * e] = f.x * scale; qv[2 * e + 1] = f.y * scale; }
  }
  const bf16* kb = k + (size_t)n * crop_stride + (size_t)hd * head_stride + sub * 8;
  const bf16* vb = v + (size_t)n * crop_stride + (size_t)hd * head_stride + sub * 8;
  float mx = -INFINITY;
#pragma unroll 8
  for (int j0 = 0; j0 < L; j0 += AD_GROUPS) {                  // uniform trip count: the shuffles below need the whole warp
    const int j = j0 + grp;
    float acc = 0.f;
    if (j < L) {
      const uint4 u = *reinterpret_cast<const uint4*>(kb + (size_t)j * ldkv);
      const bf16x2* h = reinterpret_cast<const bf16x2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = unpack2(h[e]); acc += qv[2 * e] * f.x + qv[2 * e + 1] * f.y; }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (j < L) {
      if (sub == 0) sc[j] = acc;
      mx = fmaxf(mx, acc);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < AD_WARPS; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = t; j < L; j += AD_THREADS) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < AD_WARPS; ++w) sum += red[w];
  // out[d] = sum_j p[j] v[j][d]: row group grp walks every sixteenth key, 8 dimensions per lane
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int j = grp; j < L; j += AD_GROUPS) {
    const uint4 u = *reinterpret_cast<const uint4*>(vb + (size_t)j * ldkv);
    const bf16x2* h = reinterpret_cast<const bf16x2*>(&u);
    const float pj = sc[j];
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = unpack2(h[e]); acc[2 * e] += pj * f.x; acc[2 * e + 1] += pj * f.y; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[grp][sub * 8 + e] = acc[e];
  __syncthreads();
  if (t < 64) {
    float o = 0.f;
#pragma unroll
    for (int r = 0; r < AD_GROUPS; ++r) o += part[r][t];
    out[(size_t)n * D + hd * 64 + t] = f32_to_16(o / sum);
  }
}

// ---- small elementwise kernels ---------------------------------------------------------------------------------------------
__global__ void vit_assemble_kernel(const bf16* __restrict__ patches, const bf16* __restrict__ cls, const bf16* __restrict__ pos,
                                    bf16* __restrict__ h, int n, int P, int D) {
  const long long total = (long long)n * (P + 1) * (D / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c2 = (int)(i % (D / 2));
    long long r = i / (D / 2);
    const int tok = (int)(r % (P + 1)), b = (int)(r / (P + 1));
    const float2 pv = unpack2(reinterpret_cast<const bf16x2*>(pos + (size_t)tok * D)[c2]);
    const float2 xv = tok == 0 ? unpack2(reinterpret_cast<const bf16x2*>(cls)[c2])
                               : unpack2(reinterpret_cast<const bf16x2*>(patches + ((size_t)b * P + tok - 1) * D)[c2]);
    reinterpret_cast<bf16x2*>(h + ((size_t)b * (P + 1) + tok) * D)[c2] = pack2(xv.x + pv.x, xv.y + pv.y);
  }
}

__global__ void trocr_embed_kernel(const int* __restrict__ ids, int ids_ld, int t, const bf16* __restrict__ tok, const bf16* __restrict__ pos,
                                   bf16* __restrict__ x, int n, int D, float scale, const int* __restrict__ tdev) {
  pdl_go();
  pdl_wait();
  const int b = blockIdx.x;
  if (tdev) t = *tdev;
  const int id = ids[(size_t)b * ids_ld + t];
  const bf16x2* te = reinterpret_cast<const bf16x2*>(tok + (size_t)id * D);
  const bf16x2* pe = reinterpret_cast<const bf16x2*>(pos + (size_t)(t + 2) * D);        // TrOCRLearnedPositionalEmbedding: offset 2
  bf16x2* xo = reinterpret_cast<bf16x2*>(x + (size_t)b * D);
  for (int i = threadIdx.x; i < D / 2; i += blockDim.x) {
    const float2 a = unpack2(te[i]), p = unpack2(pe[i]);
    xo[i] = pack2(a.x * scale + p.x, a.y * scale + p.y);
  }
}

__global__ void kv_append_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ cache, int n, int t, int Lcap, int D,
                                 const int* __restrict__ tdev) {
  // qkv [n][3D] (q | k | v) -> cache [n][Lcap][2D] (k | v) at position t
  pdl_go();
  pdl_wait();
  const int b = blockIdx.x;
  if (tdev) t = *tdev;
  const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)b * 3 * D + D);
  uint4* dst = reinterpret_cast<uint4*>(cache + ((size_t)b * Lcap + t) * 2 * D);
  for (int i = threadIdx.x; i < 2 * D / 8; i += blockDim.x) dst[i] = src[i];
}

// cross-attention K | V of a chunk, [n][T][2][heads][64] as the GEMM writes it -> [n][2][heads][T][64]: every (crop, head) then
// streams ONE contiguous block per decode step instead of 128-byte pieces 4 KB apart (3.3 -> ~5 TB/s in the decode attention)
__global__ void kv_to_head_major_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int T, int heads) {
  const long long total = (long long)n * T * 2 * heads * 8;            // 16-byte pieces
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    long long r = i >> 3;
    const int hd = (int)(r % heads); r /= heads;
    const int kv = (int)(r & 1); r >>= 1;
    const int tok = (int)(r % T);
    const long long b = r / T;
    out[((((b * 2 + kv) * heads + hd) * T) + tok) * 8 + c] = in[i];
  }
}

__global__ void advance_position_kernel(int* t) { pdl_go(); pdl_wait(); *t += 1; }

// greedy choice over [n][ld] fp32 logits (V valid classes); writes ids[b][t + 1]; finished sequences emit pad
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int V, int ld, int* __restrict__ ids,
                                                           int ids_ld, int t, int eos, int pad, int* __restrict__ finished,
                                                           int* __restrict__ n_finished, int* __restrict__ tdev) {
  __shared__ float bv[32];
  pdl_go();
  pdl_wait();
  if (tdev) t = *tdev;
  __shared__ int bi[32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = logits + (size_t)b * ld;
  float best = -INFINITY; int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) { const float x = row[i]; if (x > best) { best = x; idx = i; } }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, d); const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane == 0) { bv[warp] = best; bi[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -INFINITY; idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, d); const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
      if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (lane == 0) {
      const int was = finished[b];
      ids[(size_t)b * ids_ld + t + 1] = was ? pad : idx;
      if (!was && idx == eos) { finished[b] = 1; atomicAdd(n_finished, 1); }
    }
  }
}

// ---- per-crop Pillow resize to SxS, in patch order ---------------------------------------------------------------------------
// Tables (host, csrc/resize_tab.h) per crop: lo/cnt [S] and kk [S][ksize] for each axis, packed back to back; meta[c] =
// {h, w, pitch, ksize_x, ksize_y, off_x (ints into tab), off_y, tmp row offset}.  Pass 1 (horizontal, all source rows) writes
// the u8 intermediate Pillow keeps; pass 2 (vertical) normalises and scatters into [crop][patch][c*P*P + py*P + px].
struct CropMeta { int h, w, pitch, ksx, ksy, offx, offy; long long tmp_off; };

__global__ void trocr_resize_h_kernel(const uint8_t* const* __restrict__ crops, const CropMeta* __restrict__ meta, const int* __restrict__ tab,
                                      uint8_t* __restrict__ tmp, int S) {
  const int c = blockIdx.y;
  const CropMeta m = meta[c];
  const int* lo = tab + m.offx; const int* cnt = lo + S; const int* kk = cnt + S;
  const uint8_t* src = crops[c];
  uint8_t* dst = tmp + m.tmp_off;
  const long long total = (long long)m.h * S * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 3);
    const int xo = (int)((i / 3) % S);
    const int y = (int)(i / (3LL * S));
    const uint8_t* row = src + (size_t)y * m.pitch + ch;
    const int x0 = lo[xo], nx = cnt[xo];
    int acc = 1 << 21;
    for (int k = 0; k < nx; ++k) acc += kk[xo * m.ksx + k] * (int)row[(x0 + k) * 3];
    acc >>= 22;
    dst[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

__global__ void trocr_resize_v_kernel(const CropMeta* __restrict__ meta, const int* __restrict__ tab, const uint8_t* __restrict__ tmp,
                                      bf16* __restrict__ patches, int S, int P) {
  const int c = blockIdx.y;
  const CropMeta m = meta[c];
  const int* lo = tab + m.offy; const int* cnt = lo + S; const int* kk = cnt + S;
  const uint8_t* src = tmp + m.tmp_off;
  const int G = S / P;                                   // patches per side
  const int Kp = 3 * P * P;
  bf16* dst = patches + (size_t)c * G * G * Kp;
  const int total = S * S * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i % 3, xo = (i / 3) % S, yo = i / (3 * S);
    const int y0 = lo[yo], ny = cnt[yo];
    int acc = 1 << 21;
    for (int k = 0; k < ny; ++k) acc += kk[yo * m.ksy + k] * (int)src[((size_t)(y0 + k) * S + xo) * 3 + ch];
    acc >>= 22;
    acc = acc < 0 ? 0 : (acc > 255 ? 255 : acc);
    // TrOCRProcessor: BGR->RGB (text_recognizer.py:50), rescale 1/255, normalise mean 0.5 std 0.5
    const float v = ((float)acc * (1.0f / 255.0f) - 0.5f) / 0.5f;
    const int rgb = 2 - ch;
    const int gy = yo / P, py = yo % P, gx = xo / P, px = xo % P;
    dst[(size_t)(gy * G + gx) * Kp + rgb * P * P + py * P + px] = f32_to_16(v);
  }
}

// pixel_values [n][3][S][S] fp32 (already preprocessed; parity harness) -> patch order
__global__ void nchw_to_patches_kernel(const float* __restrict__ x, bf16* __restrict__ patches, int n, int S, int P) {
  const int G = S / P, Kp = 3 * P * P;
  const long long total = (long long)n * 3 * S * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % S), yo = (int)((i / S) % S), ch = (int)((i / ((long long)S * S)) % 3), b = (int)(i / (3LL * S * S));
    patches[((size_t)b * G * G + (yo / P) * G + xo / P) * Kp + ch * P * P + (yo % P) * P + xo % P] = f32_to_16(x[i]);
  }
}

inline int grid_for(long long total, int bs) {
  long long g = (total + bs - 1) / bs;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

cudaError_t layernorm_rows(const bf16* x, const float* gamma, const float* beta, bf16* y, long long rows, int C, float eps,
                           cudaStream_t s, LaunchCounter* lc) {
  if (rows <= 0) return cudaSuccess;
  if (C & 1) return cudaErrorInvalidValue;
  if (rows <= 512 && C % 8 == 0 && C <= 2048) {           // decode loop: a CTA per row
    const int threads = ((C / 8 + 31) / 32) * 32;
    cudaError_t e = launch_pdl(layernorm_row_kernel, dim3((unsigned)rows), dim3(threads), 0, s, x, gamma, beta, y, C, eps);
    if (e != cudaSuccess) return e;
  } else {
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, gamma, beta, y, rows, C, eps);
  }
  if (lc) lc->n++;
  return cudaGetLastError();
}

bool skinny_gemm_supported(int M, int N, int K) { return M >= 1 && M <= 128 && K >= 32 && K % 32 == 0 && N >= 8; }

cudaError_t skinny_gemm(const bf16* X, int ldx, const bf16* W, const float* bias, const bf16* res, int ldres, void* out, int ldo,
                        int out_f32, int M, int N, int K, int act, cudaStream_t s, LaunchCounter* lc) {
  if (M <= 0) return cudaSuccess;
  if (!skinny_gemm_supported(M, N, K)) return cudaErrorInvalidValue;
  int WN = 1;
  while (WN < 2 && (N + 8 * WN - 1) / (8 * WN) > 320) WN *= 2;     // 8 or 16 columns per CTA: wider CTAs re-read the activations per warp and thrash L1
  while (WN < 8 && K % (32 * (8 / WN)) != 0) WN *= 2;              // a K slice is whole 32-wide blocks
  while (WN > 1 && (size_t)8 * WN * (2 * K + 64) > 160 * 1024) WN /= 2;
  const size_t smem = (size_t)8 * WN * (2 * K + 64);
  if (smem > 160 * 1024) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_done;
  cudaError_t ae = once_per_device(attr_done, [] {
    cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(skinny_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(skinny_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    return e;
  });
  if (ae != cudaSuccess) return ae;
  const unsigned grid = (unsigned)((N + 8 * WN - 1) / (8 * WN));
  cudaError_t le;
  if (M <= 32) le = launch_pdl(skinny_gemm_kernel<2>, dim3(grid), dim3(256), smem, s, X, ldx, W, bias, res, ldres, out, ldo, out_f32, M, N, K, act, WN);
  else if (M <= 64) le = launch_pdl(skinny_gemm_kernel<4>, dim3(grid), dim3(256), smem, s, X, ldx, W, bias, res, ldres, out, ldo, out_f32, M, N, K, act, WN);
  else le = launch_pdl(skinny_gemm_kernel<8>, dim3(grid), dim3(256), smem, s, X, ldx, W, bias, res, ldres, out, ldo, out_f32, M, N, K, act, WN);
  if (lc) lc->n++;
  return le;
}

cudaError_t attention_enc(const bf16* qkv, bf16* out, int n, int S, int heads, float scale, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  attention_enc_kernel<<<dim3((S + 63) / 64, heads, n), 128, 0, s>>>(qkv, out, S, heads, scale * 1.4426950408889634f);
  if (lc) lc->n++;
  return cudaGetLastError();
}

cudaError_t attention_decode(const bf16* q, int ldq, const bf16* k, const bf16* v, int ldkv, int Lcap, int L, int n, int heads, float scale,
                             bf16* out, cudaStream_t s, LaunchCounter* lc, const int* tdev, int head_major) {
  if (n <= 0) return cudaSuccess;
  if (L <= 0 || L > AD_MAXL) return cudaErrorInvalidValue;
  // token-major: [n][Lcap][ldkv], head hd at column hd*64; head-major (kv_to_head_major): [n][2][heads][Lcap][64], rows of a head contiguous
  const long long crop_stride = head_major ? (long long)2 * heads * Lcap * 64 : (long long)Lcap * ldkv;
  const int head_stride = head_major ? Lcap * 64 : 64;
  cudaError_t e = launch_pdl(attention_decode_kernel, dim3(heads, n), dim3(AD_THREADS), 0, s, q, ldq, k, v, head_major ? 64 : ldkv, crop_stride,
                             head_stride, L, scale, out, heads * 64, tdev);
  if (lc) lc->n++;
  return e;
}

cudaError_t vit_assemble(const bf16* patches, const bf16* cls, const bf16* pos, bf16* h, int n, int P, int D, cudaStream_t s,
                         LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  vit_assemble_kernel<<<grid_for((long long)n * (P + 1) * (D / 2), 256), 256, 0, s>>>(patches, cls, pos, h, n, P, D);
  if (lc) lc->n++;
  return cudaGetLastError();
}

cudaError_t trocr_embed(const int* ids, int ids_ld, int t, const bf16* tok, const bf16* pos, bf16* x, int n, int D, float scale,
                        cudaStream_t s, LaunchCounter* lc, const int* tdev) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = launch_pdl(trocr_embed_kernel, dim3(n), dim3(256), 0, s, ids, ids_ld, t, tok, pos, x, n, D, scale, tdev);
  if (lc) lc->n++;
  return e;
}

cudaError_t kv_append(const bf16* qkv, bf16* cache, int n, int t, int Lcap, int D, cudaStream_t s, LaunchCounter* lc, const int* tdev) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = launch_pdl(kv_append_kernel, dim3(n), dim3(256), 0, s, qkv, cache, n, t, Lcap, D, tdev);
  if (lc) lc->n++;
  return e;
}

cudaError_t argmax_rows(const float* logits, int n, int V, int ld, int* ids, int ids_ld, int t, int eos, int pad, int* finished,
                        int* n_finished, cudaStream_t s, LaunchCounter* lc, int* tdev) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = launch_pdl(argmax_rows_kernel, dim3(n), dim3(1024), 0, s, logits, V, ld, ids, ids_ld, t, eos, pad, finished, n_finished, tdev);
  if (lc) lc->n++;
  return e;
}

cudaError_t kv_to_head_major(const bf16* in, bf16* out, int n, int T, int heads, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  kv_to_head_major_kernel<<<grid_for((long long)n * T * 2 * heads * 8, 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(in),
                                                                                         reinterpret_cast<uint4*>(out), n, T, heads);
  if (lc) lc->n++;
  return cudaGetLastError();
}

cudaError_t advance_position(int* tdev, cudaStream_t s, LaunchCounter* lc) {
  cudaError_t e = launch_pdl(advance_position_kernel, dim3(1), dim3(1), 0, s, tdev);
  if (lc) lc->n++;
  return e;
}

cudaError_t trocr_resize_patches(const uint8_t* const* crops_dev, const void* meta_dev, const int* tab_dev, uint8_t* tmp, bf16* patches,
                                 int n, int S, int P, int max_h, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  const CropMeta* meta = reinterpret_cast<const CropMeta*>(meta_dev);
  trocr_resize_h_kernel<<<dim3(grid_for((long long)max_h * S * 3, 256) < 64 ? grid_for((long long)max_h * S * 3, 256) : 64, n), 256, 0, s>>>(
      crops_dev, meta, tab_dev, tmp, S);
  trocr_resize_v_kernel<<<dim3(64, n), 256, 0, s>>>(meta, tab_dev, tmp, patches, S, P);
  if (lc) lc->n += 2;
  return cudaGetLastError();
}

cudaError_t nchw_to_patches(const float* x, bf16* patches, int n, int S, int P, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  nchw_to_patches_kernel<<<grid_for((long long)n * 3 * S * S, 256), 256, 0, s>>>(x, patches, n, S, P);
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
