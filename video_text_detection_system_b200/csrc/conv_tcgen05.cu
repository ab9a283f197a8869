// Stage 2: tcgen05 / TMEM implicit-GEMM convolution fed by TMA (sm_100a only, bf16 in, fp32 accumulate).
//
// Replaces the cuDNN convolutions the reference reaches through nn.Conv2d in DBNet (ResNet + FPN + head
// 3x3, text_detector.py:12-86) and CRNN (text_recognizer.py:16-27), plus the LSTM input projections and
// any other [pixels x Cin] x [Cin x Cout] contraction, for every layer with Cin % 64 == 0 and Cout % 64 == 0.
//
// GEMM view: M = output pixels, N = Cout, K = taps x Cin.  There is no im2col buffer and no im2col tensor
// map: activations are NHWC, so the A operand of one (tap, 64-channel chunk) for a tile of BN x BH x BW
// output pixels (product 128) is a plain 4-D TMA box {64 ch, BW, BH, BN} whose coordinates are the tile
// origin shifted by the tap; TMA's out-of-bounds zero fill IS the convolution padding.  The box lands in
// shared memory as 128 rows of 128 bytes with the 128B swizzle, which is exactly the K-major SWIZZLE_128B
// operand layout tcgen05.mma reads.  Stride-2 layers use four "parity" tensor maps (base pointer offset by
// (py,px), W and H strides doubled), so every tap is still a dense box.  Weights are [Cout][taps*Cin]
// (K-major), one 2-D box {64, BLOCK_N} per stage.
//
// CTA = 10 warps, persistent over output tiles (static stride schedule):
//   warp 0        : TMA producer (whole warp converged, one elected lane issues): full/empty mbarrier ring; one ring slot
//                   holds KPS K steps (the slot hand-over costs ~500 cycles, more than the MMAs of one narrow K step)
//   warp 1        : TMEM allocator + MMA issuer: tcgen05.mma 128 x BLOCK_N x 16, 2-4 accumulators in TMEM so the epilogue
//                   of tile i overlaps the MMAs of tile i+1; probes the next slot (mbarrier.test_wait) before issuing;
//                   tcgen05.commit releases smem slots / publishes the tile
//   warps 2..9    : epilogue (two per TMEM lane quarter): tcgen05.ld -> +bias (folded BN) -> +residual (same size, or
//                   nearest-2x upsampled = the FPN top-down add; its box arrives by TMA) -> ReLU -> bf16 / fp32 -> 4 KB
//                   128B-swizzled staging block per warp -> cp.async.bulk.tensor store (optionally 2x2 max-pooled first).
// Every mbarrier wait is bounded and traps instead of hanging the GPU.
//
// Operand delivery, by layer shape (all decided in the plan, see tc_plan_create / plan_smem):
//   * generic           : one tap-shifted 4-D box per K step, weights streamed or resident (small K, one N block);
//   * halo (3x3 s1 p1)  : ONE (8+2) x (16+2) pixel patch per tile and 64-channel chunk; the nine taps are nine start
//                         addresses of a SWIZZLE_128B descriptor whose 8-row groups are 1280 B apart; weights resident
//                         (64->64) or streamed by tap through a second ring;
//   * direct windows    : the 7x7 s2 stem reads overlapping 64-byte windows of bulk-copied input rows through a
//                         no-swizzle descriptor (row pitch 16 B);
//   * CTA pairs         : conv_tc2_kernel (halo) / conv_tc2g_kernel (generic) run the N = 256 / 128 layers with streamed
//                         weights as tcgen05.mma.cta_group::2 on clusters of two CTAs, each CTA holding half of every
//                         weight slab (the single-CTA wide tiles are bound by shared-memory bandwidth).
//
// The same warp-specialised skeleton runs four modes (template parameter MODE):
//   MODE_CONV   : the implicit-GEMM convolution described above.
//   MODE_WIN    : 3-channel stems (DBNet 7x7 s2, CRNN 3x3): activations live in a zero-bordered buffer with 4 or 8
//                 channels per pixel, and one K step is a whole filter ROW: a 5-D tensor map whose ox dimension has a
//                 16-byte stride (overlapping 64-byte windows = 8 or 4 consecutive pixels) delivers, per output
//                 pixel, the 32 bf16 that one filter row touches; 64B swizzle, K = rows x 32 (zero weights pad it).
//   MODE_DBHEAD : the DB head tail (text_detector.py:61-81 after the 3x3): per branch ConvT(64->64,k2,s2)+BN+ReLU is a
//                 [pixels x 64] x [64 x 256] GEMM (256 = 2x2 positions x 64 ch); its epilogue applies ReLU, the second
//                 ConvT (64->1, k2, s2: four 64-long dot products per position, weights in the constant bank),
//                 the optional logit bias, sigmoid and `> thr`, and writes the 4x4 output block of the pixel:
//                 probability, threshold and mask leave in one pass, no intermediate map exists.
//   MODE_LSTM   : one timestep of a BiLSTM layer for both directions: gates = h_{t-1} W_hh^T as a GEMM with rows of
//                 W_hh permuted so that a 256-column tile holds i,f,g,o of 64 hidden units; the epilogue adds the
//                 input projection, applies the cell update and writes h (bf16, next step's A operand), c and the
//                 layer output.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <stdio.h>

namespace vtd {

namespace {

using namespace tc;
constexpr int BLOCK_K = 64;          // bf16 elements = one 128-byte swizzle row
constexpr int NUM_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane quarter)
constexpr int NUM_EPI_WARPS = 8;
constexpr int HALO_PW = 10, HALO_PH = 18;                 // patch of an 8 x 16 tile of a 3x3 pad-1 conv
constexpr int HALO_BYTES = HALO_PW * HALO_PH * 128;      // 23040: one 64-channel chunk of the patch
constexpr int HALO_SLOT = 23 * 1024;
constexpr int WIN2_SLAB = 2112;       // bytes of one input row a 128-pixel tile touches: (2 * 127 + 8) px * 8 B, rounded to 16
constexpr int WIN2_SLOT = 15 * 1024;            // one ring slot = the 7 filter rows of a tile (14784 B), 1 KB granular

struct alignas(64) TcMaps {
  CUtensorMap a[4];     // activation maps; [0] only for stride 1, [py*2+px] for stride 2
  CUtensorMap b;        // weights
  CUtensorMap o;        // output (TMA-store epilogue): box = one epilogue warp's 32 pixels x 128 bytes of channels
  CUtensorMap r;        // residual (TMA-store epilogue with p.res_tma): the source pixels of that box
  CUtensorMap b4;       // halo mode with streamed weights: {64, Cout, 64-channel chunk, tap}: one box = b_taps taps of one chunk
};

enum { MODE_CONV = 0, MODE_WIN = 1, MODE_DBHEAD = 2, MODE_LSTM = 3 };

struct TcParams {
  int N, Ho, Wo, Cout, Cin;
  int KH, KW, stride, pad;
  int lw, lh;                 // log2(BW), log2(BH); BN = 128 >> (lw+lh)
  int tiles_x, tiles_y, tiles_n, n_blocks;
  int total_tiles;
  // images of this launch known only on the device (the recogniser's crop count): N = clamp(*n_dyn - n_first, 0, N), the tile
  // count follows; the grid was sized for N.  nullptr = the host values above.
  const int* n_dyn; int n_first;
  int relu, res_mode, out_f32;
  const float* bias;
  const bf16* res;
  void* out;
  int stages, bres;           // operand ring depth; 1 = all weight K-slices stay resident in shared memory
  int epi_tma;                // 1 = epilogue stages 32 px x 128 B per warp in shared memory and leaves with a TMA store
  int pool;                   // max-pool fused into the TMA-store epilogue (0 none, 1 = 2x2 s2, 2 = (2,1) s(2,1)); warp box is 16 x 2 px
  int res_tma;                // 1 = the residual of each warp's box arrives by TMA into shared memory, one item ahead
  int res_bytes;              // bytes of one residual box
  int halo;                   // 1 = 3x3 s1 p1 conv read from ONE halo patch per tile and 64-channel chunk: tile 8 x 16 px, patch 10 x 18 px
                              //     (SWIZZLE_128B as TMA stores it); the 9 taps are 9 descriptor start addresses (row shifts)
  int cta2;                   // halo == 2 on CTA pairs (conv_tc2_kernel): tcgen05.mma.cta_group::2, each CTA streams half of the weights
  int b_stages, b_taps;       // halo == 2 (weights streamed): second ring of b_stages slots, each b_taps taps of one 64-channel chunk
  int kps;                    // K steps per ring slot: one full/empty handshake (and one tcgen05.commit) per kps steps
  long long* timers;          // VTD_TIMERS builds: [grid][3 roles][total, wait, wait2] cycles
  int dbg;                    // VTD_DBG timing experiments (results are wrong): 1 no A loads, 2 no MMAs, 4 no stores
  // MODE_WIN
  int nr, sdiv;               // filter rows (= K steps), row phases (= conv stride)
  int win2;                   // 1 = direct windows: one-row tiles, the nr input rows of a tile are copied ONCE (bulk copies) and
                              //     the MMA reads the overlapping 64-byte windows in place (no-swizzle descriptor, 16-byte row pitch)
  const uint8_t* win_in;      // win2: padded input, row / image pitches in bytes
  long long win_rp, win_ip;
  // MODE_DBHEAD
  const float* logit_bias; float* prob; float* thresh; uint8_t* mask;
  // MODE_LSTM
  const float* xproj; float* cbuf; bf16* h_next; bf16* seq_out;
  int lstm_T, lstm_step, lstm_B, lstm_Bcap;
};

struct HeadConsts {            // DB head tail constants, passed in the kernel parameter (constant) bank
  float b1[2][256];            // [head][(dy,dx,c)]  folded BN shift of ConvT1
  float w2[2][64][4];          // [head][c][(dy2,dx2)]
  float b2[2];
  float thr;
  float pad_;
};
struct NoExtra { int unused; };

struct DynCount { int N, tiles_n, total_tiles; };
__device__ __forceinline__ DynCount dyn_count(const TcParams& p) {
  DynCount d{p.N, p.tiles_n, p.total_tiles};
  if (p.n_dyn) {
    int c = *p.n_dyn - p.n_first;
    c = c < 0 ? 0 : (c > p.N ? p.N : c);
    const int bnn = 128 >> (p.lw + p.lh);
    d.N = c; d.tiles_n = (c + bnn - 1) / bnn;
    d.total_tiles = p.tiles_x * p.tiles_y * d.tiles_n * p.n_blocks;
  }
  return d;
}

// -DVTD_TIMERS (dev builds only): per-role wait/total cycle counters, printed per launch by launch_tc()
#ifdef VTD_TIMERS
#define TMR_DECL long long tmr_wait = 0, tmr_wait2 = 0; const long long tmr_t0 = clock64();
#define TMR_WAIT(acc, stmt) { const long long _a = clock64(); stmt; acc += clock64() - _a; }
#define TMR_STORE(role) if (lane == 0 && p.timers) { long long* o = p.timers + ((size_t)blockIdx.x * 3 + role) * 3; \
    o[0] = clock64() - tmr_t0; o[1] = tmr_wait; o[2] = tmr_wait2; }
#else
#define TMR_DECL
#define TMR_WAIT(acc, stmt) { stmt; }
#define TMR_STORE(role)
#endif
template <int MODE> struct ExtraOf { typedef NoExtra type; };
template <> struct ExtraOf<MODE_DBHEAD> { typedef HeadConsts type; };


template <int BLOCK_N, int MODE>
struct TcCfg {
  static constexpr int ROWB = MODE == MODE_WIN ? 64 : 128;           // bytes of K per smem row
  static constexpr int A_BYTES = BLOCK_M * ROWB;
  static constexpr int B_STAGE_BYTES = BLOCK_N * ROWB;
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = MODE == MODE_WIN ? 12 : ((BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8));
  static constexpr int MAX_STAGES = 16;
  static constexpr int TAIL_BYTES = 8 * (2 * MAX_STAGES + 18) + 16 + (MODE == MODE_DBHEAD ? 4096 : 0);   // barriers, TMEM slot, head consts
  // dynamic shared memory for a given ring depth / resident-weight size
  static constexpr int smem_bytes(int stages, int bres_bytes) {
    return stages * (bres_bytes ? A_BYTES : STAGE_BYTES) + bres_bytes + 1024 /*alignment slack*/ + TAIL_BYTES;
  }
  static constexpr int ACC = 512 / BLOCK_N > 4 ? 4 : 512 / BLOCK_N;        // accumulator buffers in TMEM: 4 / 4 / 2
  static constexpr int TMEM_COLS = ACC * BLOCK_N;                          // 256 / 512 / 512: powers of two

};


// No-swizzle K-major descriptor (core matrix = 8 rows x 16 bytes, rows 16 bytes apart): lbo = byte stride between core
// matrices along K, sbo = along M.  With lbo = 16, sbo = 128 row m starts 16 bytes after row m-1 and rows OVERLAP: exactly
// the im2col of a convolution whose consecutive outputs are 16 bytes apart in a padded NHWC row
// (profiles/micro/umma_nosw.cu checks the hardware reads it that way).
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// SWIZZLE_128B K-major descriptor whose 8-row groups are `sbo` bytes apart and whose start may be any 128-byte row of a
// swizzled region: the XOR pattern is a function of the absolute shared-memory address (profiles/micro/umma_shift.cu:
// exact for start rows 1, 11, 21 with sbo = 1280 and the base-offset field left 0).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// ---- epilogues ---------------------------------------------------------------------------------------------

// MODE_CONV / MODE_WIN: bias + residual + ReLU -> NHWC bf16 / fp32.  Two warps share a TMEM lane quarter and split
// the tile's columns (`half`).  All residual vectors of the warp's columns are requested BEFORE the accumulator
// barrier is waited on, so their DRAM/L2 latency overlaps the MMAs instead of serialising per 32-column chunk.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_conv(const TcParams& p, uint32_t tmem_acc, int q, int half, bool valid,
                                              size_t opix, size_t rpix, int nb, uint32_t tfull_bar, uint32_t parity) {
  constexpr int NCH = BLOCK_N / 64;                  // 32-column chunks per warp
  const int cbase = half * (BLOCK_N / 2);
  uint4 rv[NCH][4];
  const bool has_res = p.res_mode != RES_NONE && !(VTD_DBG_BITS(p) & 12);
  if (VTD_DBG_BITS(p) & 4) valid = false;
  if (has_res && valid) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix * p.Cout + nb * BLOCK_N + cbase);
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) rv[i][j] = __ldg(rp + i * 4 + j);
  }
  mbar_wait(tfull_bar, parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c0 = cbase + i * 32;
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (valid) {
      const int co = nb * BLOCK_N + c0;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co + j));
          f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
        }
      }
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bf16x2* h = reinterpret_cast<const bf16x2*>(&rv[i][j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 r2 = unpack2(h[e]);
            f[j * 8 + e * 2] += r2.x; f[j * 8 + e * 2 + 1] += r2.y;
          }
        }
      }
      if (p.relu == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (p.relu == 2) {                       // exact GELU (transformer MLPs): x/2 (1 + erf(x / sqrt 2))
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f));
      }
      if (p.out_f32) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * p.Cout + co);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      } else {
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + opix * p.Cout + co);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          bf16x2 h0 = pack2(f[8 * j], f[8 * j + 1]);
          bf16x2 h1 = pack2(f[8 * j + 2], f[8 * j + 3]);
          bf16x2 h2 = pack2(f[8 * j + 4], f[8 * j + 5]);
          bf16x2 h3 = pack2(f[8 * j + 6], f[8 * j + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          op[j] = u;
        }
      }
    }
  }
}

// MODE_CONV / MODE_WIN with p.epi_tma: the same arithmetic, but the tile leaves through shared memory and TMA stores.
// Per-thread stores write 16 bytes per lane at a pixel stride (32 sectors touched per warp instruction, half of each
// used); here a warp packs its 32 pixels x 128 bytes of channels (64 bf16 or 32 fp32: one "group") into a 4 KB
// 128B-swizzled staging block and one lane issues cp.async.bulk.tensor: full 128-byte rows, out-of-range pixels are
// clipped by the tensor map, and the store drains while the warp converts the next group.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_group_tma(const TcParams& p, const CUtensorMap* omap, uint32_t stg, uint32_t pstg,
                                                   uint32_t tmem_acc, int q, int lane, int g, bool use_res,
                                                   const uint4 (&rv)[8], int nb, int cx, int cy, int cn) {
  const int gcols = p.out_f32 ? 32 : 64;                 // accumulator columns per 128-byte group
  const uint32_t row = stg + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  const int c0 = g * gcols;                              // first column of the group inside the tile
  const int co = nb * BLOCK_N + c0;
  uint4 o[8];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    if (p.out_f32 && hh == 1) break;
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + hh * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co + hh * 32 + j));
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    }
    if (use_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bf16x2* h = reinterpret_cast<const bf16x2*>(&rv[hh * 4 + j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 r2 = unpack2(h[e]);
          f[j * 8 + e * 2] += r2.x; f[j * 8 + e * 2 + 1] += r2.y;
        }
      }
    }
    if (p.relu == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    } else if (p.relu == 2) {                         // exact GELU (transformer MLPs)
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f));
    }
    if (p.out_f32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                          __float_as_uint(f[4 * j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf16x2 h0 = pack2(f[8 * j], f[8 * j + 1]);
        bf16x2 h1 = pack2(f[8 * j + 2], f[8 * j + 3]);
        bf16x2 h2 = pack2(f[8 * j + 4], f[8 * j + 5]);
        bf16x2 h3 = pack2(f[8 * j + 6], f[8 * j + 7]);
        o[hh * 4 + j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      }
    }
  }
  if (VTD_DBG_BITS(p) & 4) return;
  // the previous store of this warp must have finished READING the staging block before it is overwritten
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)j ^ sw) << 4)), "r"(o[j].x), "r"(o[j].y),
                 "r"(o[j].z), "r"(o[j].w) : "memory");
  if (p.pool) {
    // fused max-pool (nn.MaxPool2d after conv+BN+ReLU, text_recognizer.py:17-23): the warp's box is 16 x 2 pixels (8 x 4 in
    // the halo tiles), so every pooling window lies inside it.  Each lane reduces a slice of one pooled pixel from the
    // staged rows with packed bf16 max (max commutes with the bf16 rounding already applied) and only the pooled box goes
    // to HBM.  Staged row of box pixel (x, y) = y * bw + x; pooled pixel pp = py * (pooled box width) + px.
    __syncwarp();
    const bool p22 = p.pool == 1;
    const int bw = 1 << p.lw;                                   // box width: 16 or 8
    const int pp = p22 ? lane >> 2 : lane >> 1;                 // pooled pixel of the box: 8 (2x2) or 16 ((2,1))
    const int c_first = p22 ? (lane & 3) * 2 : (lane & 1) * 4;  // first 16-byte chunk of this lane's slice
    const int nch = p22 ? 2 : 4;
    const int pbw = p22 ? bw >> 1 : bw;
    const int ppx = pp & (pbw - 1), ppy = pp / pbw;
    const int r00 = 2 * ppy * bw + (p22 ? 2 * ppx : ppx);       // top-left source row (pixel) of the window
    const uint32_t prow = pstg + (uint32_t)pp * 128u;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      if (ci >= nch) break;
      const uint32_t c = (uint32_t)(c_first + ci);
      uint4 m4;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        if (!p22 && (w & 1)) continue;                          // (2,1): the pixel and the one below it only
        const uint32_t r = (uint32_t)(r00 + (w & 1) + (w >> 1) * bw);
        uint4 t;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                     : "r"(stg + r * 128u + ((c ^ (r & 7u)) << 4)) : "memory");
        if (w == 0) m4 = t;
        else {
          bf16x2 a, b;
#define VTD_HMAX2(dst, src) a = *reinterpret_cast<bf16x2*>(&dst); b = *reinterpret_cast<bf16x2*>(&src); \
          a = __hmax2(a, b); dst = *reinterpret_cast<uint32_t*>(&a);
          VTD_HMAX2(m4.x, t.x) VTD_HMAX2(m4.y, t.y) VTD_HMAX2(m4.z, t.z) VTD_HMAX2(m4.w, t.w)
#undef VTD_HMAX2
        }
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((c ^ ((uint32_t)pp & 7u)) << 4)), "r"(m4.x), "r"(m4.y),
                   "r"(m4.z), "r"(m4.w) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(omap), "r"(pstg),
                   "r"(co), "r"(p22 ? cx >> 1 : cx), "r"(cy >> 1), "r"(cn) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    return;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(omap), "r"(stg),
                 "r"(co), "r"(cx), "r"(cy), "r"(cn) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// MODE_DBHEAD: ReLU(ConvT1) -> ConvT2 -> (+logit bias) -> sigmoid -> 4x4 block of prob / thresh (+ mask).
// hs points at this head's constants in shared memory: b1[256], then w2[64][4].  Warp `half` owns ConvT1 positions
// dy = half (g = 2*half, 2*half+1), i.e. output rows 2*half and 2*half+1 of the pixel's 4x4 block.  The loops over
// (g, 32-channel chunk) stay rolled: the body is small enough for the instruction cache (a fully unrolled version
// with immediate constants was instruction-fetch bound, profiles/r01_ncu_notes.md).
__device__ __forceinline__ void epilogue_dbhead(const TcParams& p, const float* __restrict__ hs, float b2, float thr,
                                                int head, uint32_t tmem_acc, int q, int half, bool valid, int n, int oy,
                                                int ox) {
  float o[8];                                       // [dx][(dy2,dx2)]
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = b2;
#pragma unroll 2
  for (int it = 0; it < 4; ++it) {                  // (dx, 32-channel half)
    const int dx = it >> 1, ch = it & 1;
    const int g = half * 2 + dx;
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float* b1 = hs + g * 64 + ch * 32;
    const float4* w2 = reinterpret_cast<const float4*>(hs + 256) + ch * 32;
    // packed fp32 (FADD2 / FFMA2, sm_100): the same IEEE operations in the same order, two lanes per instruction
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float2 h = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                            *reinterpret_cast<const float2*>(b1 + j));
      h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f);
      const float4 w0 = w2[j], w1 = w2[j + 1];
      a01 = __ffma2_rn(make_float2(h.x, h.x), make_float2(w0.x, w0.y), a01);
      a23 = __ffma2_rn(make_float2(h.x, h.x), make_float2(w0.z, w0.w), a23);
      a01 = __ffma2_rn(make_float2(h.y, h.y), make_float2(w1.x, w1.y), a01);
      a23 = __ffma2_rn(make_float2(h.y, h.y), make_float2(w1.z, w1.w), a23);
    }
    if (dx == 0) { o[0] += a01.x; o[1] += a01.y; o[2] += a23.x; o[3] += a23.y; }
    else { o[4] += a01.x; o[5] += a01.y; o[6] += a23.x; o[7] += a23.y; }
  }
  if (!valid) return;
  const int Wd = 4 * p.Wo, Hd = 4 * p.Ho;
  float* __restrict__ outp = head == 0 ? p.prob : p.thresh;
#pragma unroll
  for (int dy2 = 0; dy2 < 2; ++dy2) {
    const int r = half * 2 + dy2;
    float v0 = o[dy2 * 2 + 0], v1 = o[dy2 * 2 + 1], v2 = o[4 + dy2 * 2 + 0], v3 = o[4 + dy2 * 2 + 1];
    const size_t oidx = ((size_t)n * Hd + 4 * oy + r) * Wd + 4 * ox;
    if (head == 0 && p.logit_bias) {
      float4 lb = __ldg(reinterpret_cast<const float4*>(p.logit_bias + oidx));
      v0 += lb.x; v1 += lb.y; v2 += lb.z; v3 += lb.w;
    }
    v0 = 1.0f / (1.0f + expf(-v0)); v1 = 1.0f / (1.0f + expf(-v1));
    v2 = 1.0f / (1.0f + expf(-v2)); v3 = 1.0f / (1.0f + expf(-v3));
    *reinterpret_cast<float4*>(outp + oidx) = make_float4(v0, v1, v2, v3);
    if (head == 0) {
      uint32_t m = (v0 > thr ? 1u : 0u) | (v1 > thr ? 0x100u : 0u) | (v2 > thr ? 0x10000u : 0u) |
                   (v3 > thr ? 0x1000000u : 0u);
      *reinterpret_cast<uint32_t*>(p.mask + oidx) = m;
    }
  }
}

// pull this thread's input-projection and cell-state lines towards L2 while the MMAs of the step run
__device__ __forceinline__ void lstm_prefetch(const TcParams& p, int half, int b, int nb) {
  if (b >= p.lstm_B) return;
  const int dir = nb >> 2, jt = nb & 3;
  const int t = dir == 0 ? p.lstm_step : p.lstm_T - 1 - p.lstm_step;
  const float* xp = p.xproj + (((size_t)b * p.lstm_T + t) * 2 + dir) * 1024 + jt * 256 + half * 32;
  const float* cs = p.cbuf + ((size_t)dir * p.lstm_Bcap + b) * 256 + jt * 64 + half * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + g * 64));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(cs));
}

// MODE_LSTM: gates -> cell update.  Tile columns = [i(64) | f(64) | g(64) | o(64)] of hidden units jt*64..+63.
__device__ __forceinline__ void epilogue_lstm(const TcParams& p, uint32_t tmem_acc, int q, int half, int b, int nb) {
  const int dir = nb >> 2, jt = nb & 3;
  const bool valid = b < p.lstm_B;
  const int t = dir == 0 ? p.lstm_step : p.lstm_T - 1 - p.lstm_step;
  const int bb = valid ? b : 0;
  const float* __restrict__ xp = p.xproj + (((size_t)bb * p.lstm_T + t) * 2 + dir) * 1024 + jt * 256;
  float* __restrict__ cs = p.cbuf + ((size_t)dir * p.lstm_Bcap + bb) * 256 + jt * 64;
  bf16* __restrict__ hn = p.h_next + ((size_t)dir * p.lstm_Bcap + bb) * 256 + jt * 64;
  bf16* __restrict__ so = p.seq_out + ((size_t)bb * p.lstm_T + t) * 512 + dir * 256 + jt * 64;
#pragma unroll 1
  for (int qq = half * 2; qq < half * 2 + 2; ++qq) {   // 16 hidden units at a time; 32 units per warp
    uint32_t vi[16], vf[16], vg[16], vo[16];
    const uint32_t base = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(qq * 16);
    tmem_ld16(base, vi); tmem_ld16(base + 64, vf); tmem_ld16(base + 128, vg); tmem_ld16(base + 192, vo);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (valid) {
      float hv[16];
#pragma unroll
      for (int u4 = 0; u4 < 16; u4 += 4) {
        const int u = qq * 16 + u4;
        float4 xi = __ldg(reinterpret_cast<const float4*>(xp + u));
        float4 xf = __ldg(reinterpret_cast<const float4*>(xp + 64 + u));
        float4 xg = __ldg(reinterpret_cast<const float4*>(xp + 128 + u));
        float4 xo = __ldg(reinterpret_cast<const float4*>(xp + 192 + u));
        float4 c4 = *reinterpret_cast<const float4*>(cs + u);
        const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
        const float xgv[4] = {xg.x, xg.y, xg.z, xg.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
        float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float gi = __uint_as_float(vi[u4 + e]) + xiv[e], gf = __uint_as_float(vf[u4 + e]) + xfv[e];
          float gg = __uint_as_float(vg[u4 + e]) + xgv[e], go = __uint_as_float(vo[u4 + e]) + xov[e];
          float cn = fast_sigmoid(gf) * cv[e] + fast_sigmoid(gi) * fast_tanh(gg);
          cv[e] = cn;
          hv[u4 + e] = fast_sigmoid(go) * fast_tanh(cn);
        }
        *reinterpret_cast<float4*>(cs + u) = make_float4(cv[0], cv[1], cv[2], cv[3]);
      }
      uint4 w0, w1;
      bf16x2 h2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) h2[e] = pack2(hv[2 * e], hv[2 * e + 1]);
      w0.x = *reinterpret_cast<uint32_t*>(&h2[0]); w0.y = *reinterpret_cast<uint32_t*>(&h2[1]);
      w0.z = *reinterpret_cast<uint32_t*>(&h2[2]); w0.w = *reinterpret_cast<uint32_t*>(&h2[3]);
      w1.x = *reinterpret_cast<uint32_t*>(&h2[4]); w1.y = *reinterpret_cast<uint32_t*>(&h2[5]);
      w1.z = *reinterpret_cast<uint32_t*>(&h2[6]); w1.w = *reinterpret_cast<uint32_t*>(&h2[7]);
      uint4* hp = reinterpret_cast<uint4*>(hn + qq * 16);
      hp[0] = w0; hp[1] = w1;
      uint4* sp = reinterpret_cast<uint4*>(so + qq * 16);
      sp[0] = w0; sp[1] = w1;
    }
  }
}

template <int BLOCK_N, int MODE, int KPS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p,
               const __grid_constant__ typename ExtraOf<MODE>::type ex) {
  using Cfg = TcCfg<BLOCK_N, MODE>;
  constexpr int ROWB = Cfg::ROWB;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand ring (swizzle atoms), barriers after it
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  const int stages = p.stages;
  constexpr int kps = KPS;                       // K steps per ring slot (host: ksteps % KPS == 0)
  const uint32_t step_bytes = p.bres ? Cfg::A_BYTES : Cfg::STAGE_BYTES;      // bytes one K step brings in
  const uint32_t stage_bytes = (MODE == MODE_WIN && p.win2) ? (uint32_t)WIN2_SLOT
                               : (MODE == MODE_CONV && p.halo) ? (uint32_t)HALO_SLOT : (uint32_t)kps * step_bytes;   // slot = kps A slabs, then kps B slabs
  const int ksteps_all = MODE == MODE_WIN ? p.nr : p.KH * p.KW * (p.Cin / BLOCK_K);
  const uint32_t bslot_bytes = (MODE == MODE_CONV && p.halo == 2) ? (uint32_t)p.b_taps * Cfg::B_STAGE_BYTES : 0u;
  const uint32_t bring = ring + stages * stage_bytes;                       // halo == 2: the weight ring follows the patch ring
  const uint32_t bres0 = bring + (uint32_t)(MODE == MODE_CONV && p.halo == 2 ? p.b_stages : 0) * bslot_bytes;   // resident weight K-slices (if p.bres)
  const uint32_t bres_bytes = p.bres ? (uint32_t)ksteps_all * Cfg::B_STAGE_BYTES : 0u;
  const uint32_t stg_bytes = (p.epi_tma ? (uint32_t)NUM_EPI_WARPS * 4096u : 0u) + (p.res_tma ? (uint32_t)NUM_EPI_WARPS * 4096u : 0u) +
                             (p.pool ? (uint32_t)NUM_EPI_WARPS * 2048u : 0u);
  const uint32_t stg0 = bres0 + bres_bytes;                                  // 4 KB of store staging per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + (bres0 - ring) + bres_bytes + stg_bytes);
  const uint32_t full0 = smem_u32(bars);                       // [MAX_STAGES]
  const uint32_t empty0 = full0 + 8 * Cfg::MAX_STAGES;         // [MAX_STAGES]
  const uint32_t tfull0 = empty0 + 8 * Cfg::MAX_STAGES;        // [4]
  const uint32_t tempty0 = tfull0 + 32;                        // [4]
  const uint32_t bfull = tempty0 + 32;                         // [1] resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::MAX_STAGES + 9);
  const uint32_t rbar0 = bfull + 16;                           // [8] residual box landed, one per epilogue warp
  float* head_s = reinterpret_cast<float*>(bars + 2 * Cfg::MAX_STAGES + 18);   // MODE_DBHEAD: [2 heads][256 b1 + 256 w2]
  if constexpr (MODE == MODE_DBHEAD) {
    for (int i = threadIdx.x; i < 2 * 512; i += NUM_THREADS) {
      const int hd = i >> 9, r = i & 511;
      head_s[i] = r < 256 ? ex.b1[hd][r] : (&ex.w2[hd][0][0])[r - 256];
    }
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    if (MODE == MODE_CONV && p.halo == 2)
      for (int i = 0; i < p.b_stages; ++i) { mbar_init(full0 + 8 * (8 + i), 1); mbar_init(empty0 + 8 * (8 + i), 1); }
    // BLOCK_N = 64 with the TMA-store epilogue: the two halves of the epilogue take alternate tiles (4 arrivals each)
    const uint32_t epi_arrivals = (BLOCK_N == 64 && p.epi_tma) ? NUM_EPI_WARPS / 2 : NUM_EPI_WARPS;
    for (int i = 0; i < Cfg::ACC; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, epi_arrivals); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.o) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.r) : "memory");
    mbar_init(bfull, 1);
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(rbar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream; global
  // memory written by that kernel is only touched below.  (No-ops when launched without the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int BW = 1 << p.lw, BH = 1 << p.lh;
  const int BNt = BLOCK_M >> (p.lw + p.lh);
  constexpr int acc_n = Cfg::ACC;
  const int kchunks = p.Cin / BLOCK_K;
  const int ksteps = MODE == MODE_WIN ? p.nr : ((MODE == MODE_CONV && p.halo) ? kchunks : p.KH * p.KW * kchunks);   // ring slots x kps per tile
  const DynCount dyn = dyn_count(p);
  const int total_tiles = dyn.total_tiles;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    // One ring slot holds kps K steps: the per-slot handshake (empty wait, expect_tx, the consumer's full wait and
    // tcgen05.commit) costs a few hundred cycles whatever the slot holds (profiles/micro/handshake.cu), which is more
    // than the MMAs of ONE K step take on the narrow layers.
    int stage = 0; uint32_t phase = 0;
    TMR_DECL
    if (p.bres && (int)blockIdx.x < total_tiles) {      // weights once per CTA (n_blocks == 1)
      if (elect_one()) {
        mbar_expect_tx(bfull, bres_bytes);
        for (int ks = 0; ks < ksteps_all; ++ks)
          tma_load_2d(bres0 + ks * Cfg::B_STAGE_BYTES, &maps.b, bfull, ks * (ROWB / 2), 0);
      }
      __syncwarp();
    }
    if (MODE == MODE_CONV && p.halo == 2) {
      // halo mode with streamed weights: two rings.  Per tile and 64-channel chunk ONE patch (patch ring) and 9 / b_taps weight
      // boxes of b_taps taps each (weight ring): 4 + 12 TMAs per tile for 256 -> 128 instead of 72, and the activation
      // crosses L2 -> SM once instead of nine times.
      int bs = 0; uint32_t bph = 0;
      const int tgroups = 9 / p.b_taps;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int x0 = tx * BW, y0 = ty * BH, n0 = t * BNt;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full0 + 8 * stage, (uint32_t)HALO_BYTES);
            tma_load_4d(ring + stage * stage_bytes, &maps.a[1], full0 + 8 * stage, kc * BLOCK_K, x0 - 1, y0 - 1, n0);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
          for (int tg = 0; tg < tgroups; ++tg) {
            mbar_wait(empty0 + 8 * (8 + bs), bph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full0 + 8 * (8 + bs), bslot_bytes);
              tma_load_4d(bring + bs * bslot_bytes, &maps.b4, full0 + 8 * (8 + bs), 0, nb * BLOCK_N, kc, tg * p.b_taps);
            }
            __syncwarp();
            if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
          }
        }
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile;
      const int nb = t % p.n_blocks; t /= p.n_blocks;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int x0 = tx * BW, y0 = ty * BH, n0 = t * BNt;
      int r = 0, sx = 0, kc = 0;                          // filter row, filter column, 64-channel chunk of the next K step
      for (int ks0 = 0; ks0 < ksteps; ks0 += kps) {
        constexpr int nk = kps;
        TMR_WAIT(tmr_wait, mbar_wait(empty0 + 8 * stage, phase ^ 1))
        if (elect_one()) {
          const uint32_t sa = ring + stage * stage_bytes;
          const uint32_t sb = sa + (uint32_t)kps * Cfg::A_BYTES;
          const uint32_t fb = full0 + 8 * stage;
          if (MODE == MODE_CONV && p.halo) {
            // one box per 64-channel chunk: the tile's pixels plus a 1-pixel ring (out-of-range rows / columns are zero
            // filled = the padding); every tap reads it in place, so the activation crosses L2 -> SM once, not 9 times
            mbar_expect_tx(fb, (uint32_t)HALO_BYTES);
            tma_load_4d(sa, &maps.a[1], fb, ks0 * BLOCK_K, x0 - 1, y0 - 1, n0);
          } else if (MODE == MODE_WIN && p.win2) {
            // the nr input rows under this one-row tile, each copied once: 7 x 2112 B instead of 7 boxes of 128 overlapping
            // 64-byte windows (57 KB; the windowed TMA ran at ~3.8 cycles per 64-byte row and bound the kernel)
            mbar_expect_tx(fb, (uint32_t)nk * WIN2_SLAB);
            const uint8_t* src = p.win_in + (size_t)n0 * p.win_ip + (size_t)(y0 * p.sdiv) * p.win_rp + (size_t)x0 * 16;
#pragma unroll
            for (int j = 0; j < nk; ++j)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(sa + j * WIN2_SLAB), "l"(src + (size_t)j * p.win_rp), "r"((uint32_t)WIN2_SLAB), "r"(fb) : "memory");
          } else {
          if (VTD_DBG_BITS(p) & 1) {
            if (p.bres) mbar_arrive(fb); else mbar_expect_tx(fb, (uint32_t)nk * Cfg::B_STAGE_BYTES);
          } else {
            mbar_expect_tx(fb, (uint32_t)nk * step_bytes);
          }
#pragma unroll
          for (int j = 0; j < nk; ++j) {
            if (MODE == MODE_WIN) {
              if (!(VTD_DBG_BITS(p) & 1)) tma_load_5d(sa + j * Cfg::A_BYTES, &maps.a[0], fb, 0, x0, r % p.sdiv, y0 + r / p.sdiv, n0);
              if (!p.bres) tma_load_2d(sb + j * Cfg::B_STAGE_BYTES, &maps.b, fb, r * 32, nb * BLOCK_N);
              ++r;
            } else {
              int mi = 0, xo, yo, coff = 0;
              if (MODE == MODE_DBHEAD) { xo = yo = 0; coff = nb * 64; }
              else if (MODE == MODE_LSTM) { xo = yo = 0; mi = nb >> 2; }
              else if (p.stride == 1) { xo = sx - p.pad; yo = r - p.pad; }
              else {
                const int tyy = r - p.pad, txx = sx - p.pad;
                const int py = tyy & 1, px = txx & 1;
                mi = py * 2 + px; yo = (tyy - py) / 2; xo = (txx - px) / 2;
              }
              if (!(VTD_DBG_BITS(p) & 1)) tma_load_4d(sa + j * Cfg::A_BYTES, &maps.a[mi], fb, coff + kc * BLOCK_K, x0 + xo, y0 + yo, n0);
              if (!p.bres)
                tma_load_2d(sb + j * Cfg::B_STAGE_BYTES, &maps.b, fb, (r * p.KW + sx) * p.Cin + kc * BLOCK_K, nb * BLOCK_N);
              if (++kc == kchunks) { kc = 0; if (++sx == p.KW) { sx = 0; ++r; } }
            }
          }
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    TMR_STORE(0)
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    const uint32_t idesc = umma_idesc(BLOCK_N);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    TMR_DECL
    if (p.bres && (int)blockIdx.x < total_tiles) mbar_wait(bfull, 0);
    bool ready = false;                                 // the slot about to be consumed is already known to be full
    if (MODE == MODE_CONV && p.halo == 2) {
      int bs = 0; uint32_t bph = 0;
      const int tgroups = 9 / p.b_taps;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * as, aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(full0 + 8 * stage, phase);
          const uint32_t sa = ring + stage * stage_bytes;
          int fr = 0, fs = 0;                               // filter row / column of the next tap
          for (int tg = 0; tg < tgroups; ++tg) {
            if (!ready) mbar_wait(full0 + 8 * (8 + bs), bph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int nbs = bs + 1 == p.b_stages ? 0 : bs + 1;
            const uint32_t nbph = bs + 1 == p.b_stages ? bph ^ 1 : bph;
            const bool probe = mbar_test(full0 + 8 * (8 + nbs), nbph);      // next weight slot, looked at before issuing
            if (elect_one()) {
              const uint32_t sbb = bring + bs * bslot_bytes;
              const uint64_t ad0 = umma_desc_sw128(sa + (uint32_t)(fr * HALO_PW + fs) * 128u, HALO_PW * 128u);
              const uint64_t bd0 = umma_desc<ROWB>(sbb);
              if (p.b_taps == 3) {                            // one filter row per slot: taps are 128 B apart in the patch
#pragma unroll
                for (int tt = 0; tt < 3; ++tt)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16(d_tmem, ad0 + (uint64_t)(tt * 8 + k * 2), bd0 + (uint64_t)(tt * (Cfg::B_STAGE_BYTES >> 4) + k * 2), idesc,
                             (kc | tg | tt | k) ? 1u : 0u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (kc | tg | k) ? 1u : 0u);
              }
              umma_commit(empty0 + 8 * (8 + bs));                       // weight slot free when these MMAs have read it
              if (tg == tgroups - 1) {
                umma_commit(empty0 + 8 * stage);                        // ... and so is the patch after its last tap
                if (kc == kchunks - 1) umma_commit(tfull0 + 8 * as);    // accumulator complete
              }
            }
            ready = __any_sync(0xffffffffu, probe);
            if (p.b_taps == 3) ++fr; else if (++fs == 3) { fs = 0; ++fr; }
            bs = nbs; bph = nbph;
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (++as == acc_n) { as = 0; aphase ^= 1; }
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      TMR_WAIT(tmr_wait2, mbar_wait(tempty0 + 8 * as, aphase ^ 1))
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
      for (int ks0 = 0; ks0 < ksteps; ks0 += kps) {
        constexpr int nk = kps;
        if (!ready) { TMR_WAIT(tmr_wait, mbar_wait(full0 + 8 * stage, phase)) }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // look one slot ahead before issuing (test_wait never suspends): the probe's round trip to the barrier overlaps
        // the MMA issue below.  A/B on one box: -2..5 % on the N = 256 layers; the same trick in the producer, a second
        // producer warp, and warp-uniform producer bookkeeping measured neutral or worse and were dropped.
        const int nstage = stage + 1 == stages ? 0 : stage + 1;
        const uint32_t nphase = stage + 1 == stages ? phase ^ 1 : phase;
        const bool probe = mbar_test(full0 + 8 * nstage, nphase);
        if (elect_one()) {
          const uint32_t sa = ring + stage * stage_bytes;
          const uint32_t sb = sa + (uint32_t)kps * Cfg::A_BYTES;
          if (MODE == MODE_CONV && p.halo) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint64_t ad = umma_desc_sw128(sa + (uint32_t)((t / 3) * HALO_PW + (t % 3)) * 128u, HALO_PW * 128u);
              const uint64_t bd = umma_desc<ROWB>(bres0 + (uint32_t)(t * kchunks + ks0) * Cfg::B_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ks0 | t | k) ? 1u : 0u);
            }
          } else if (!(VTD_DBG_BITS(p) & 2)) {
#pragma unroll
            for (int j = 0; j < nk; ++j) {
              const uint64_t ad = (MODE == MODE_WIN && p.win2) ? umma_desc_nosw(sa + j * WIN2_SLAB, 16, 128)
                                                               : umma_desc<ROWB>(sa + j * Cfg::A_BYTES);
              const uint64_t bd = umma_desc<ROWB>(p.bres ? bres0 + (ks0 + j) * Cfg::B_STAGE_BYTES : sb + j * Cfg::B_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < ROWB / 32; ++k)
                umma_f16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ks0 | j | k) ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8 * stage);            // frees the slot when these MMAs have read it
          if (ks0 + nk >= ksteps) umma_commit(tfull0 + 8 * as);   // accumulator complete
        }
        ready = __any_sync(0xffffffffu, probe);       // a completed phase stays completed: any lane's "yes" holds for all
        stage = nstage; phase = nphase;
      }
      if (++as == acc_n) { as = 0; aphase ^= 1; }
    }
    TMR_STORE(1)
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                   // which half of the tile's columns / units / rows
    const int m = q * 32 + lane;                        // row of the tile = pixel
    const int xx = m & (BW - 1), yy = (m >> p.lw) & (BH - 1), nn = m >> (p.lw + p.lh);
    TMR_DECL
    if ((MODE == MODE_CONV || MODE == MODE_WIN) && p.epi_tma) {
      // ---- TMA-store epilogue.  Work items = (tile, 128-byte channel group); N = 64 tiles have one group, so the two
      // halves take alternate tiles; wider tiles split their groups between the halves.  The residual of the NEXT work
      // item (next group, or the first group of this warp's next tile) is requested before the current one is
      // converted, so its L2/DRAM latency never sits between an accumulator and its store.
      const int step = (BLOCK_N == 64 ? 2 : 1) * (int)gridDim.x;
      const int groups = (p.out_f32 ? BLOCK_N * 4 : BLOCK_N * 2) / 128;
      const int g0 = BLOCK_N == 64 ? 0 : half * (groups / 2), g1 = BLOCK_N == 64 ? groups : g0 + groups / 2;
      const bool use_res = p.res_mode != RES_NONE && !(VTD_DBG_BITS(p) & 12);
      const uint32_t stg = stg0 + (uint32_t)(warp - 2) * 4096u;
      const uint32_t pstg = stg0 + (uint32_t)NUM_EPI_WARPS * (4096u + (p.res_tma ? 4096u : 0u)) + (uint32_t)(warp - 2) * 2048u;
      const int m0 = q * 32;                              // first pixel of this warp inside the tile
      const int gcols = p.out_f32 ? 32 : 64;
      // residual: per-lane loads (16 B per lane at a pixel stride) cost one L1 tag lookup per lane and instruction; with
      // p.res_tma the source pixels of the warp's box arrive by TMA (requested one work item ahead) and each lane reads
      // its 128-byte row from shared memory.
      const uint32_t rbuf = stg0 + (uint32_t)NUM_EPI_WARPS * 4096u + (uint32_t)(warp - 2) * 4096u;
      const uint32_t rbar = rbar0 + 8 * (uint32_t)(warp - 2);
      const int up = p.res_mode == RES_UP2 ? 1 : 0;
      const int sw_ = BW < 32 ? BW : 32, sh_ = BH < 32 / sw_ ? BH : 32 / sw_;           // the warp's box: sw_ x sh_ x sn_ pixels
      const int lx = lane & (sw_ - 1), ly = (lane / sw_) & (sh_ - 1), ln = lane / (sw_ * sh_);
      const int rw = up ? (sw_ > 1 ? sw_ >> 1 : 1) : sw_, rh = up ? (sh_ > 1 ? sh_ >> 1 : 1) : sh_;
      const int rrow = (ln * rh + (up ? (sh_ > 1 ? ly >> 1 : 0) : ly)) * rw + (up ? lx >> 1 : lx);   // my row of the residual box
      const uint32_t raddr = rbuf + (uint32_t)rrow * 128u;
      const uint32_t rsw = (uint32_t)(rrow & 7);
      uint32_t rphase = 0;
      auto res_issue = [&](int tile_, int g_) {
        int t = tile_;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int cx = tx * BW + (m0 & (BW - 1)), cy = ty * BH + ((m0 >> p.lw) & (BH - 1)), cn = t * BNt + (m0 >> (p.lw + p.lh));
        if (lane == 0) {
          mbar_expect_tx(rbar, (uint32_t)p.res_bytes);
          tma_load_4d(rbuf, &maps.r, rbar, nb * BLOCK_N + g_ * gcols, cx >> up, cy >> up, cn);
        }
      };
      auto res_load = [&](int tile_, int g_, uint4 (&r)[8]) {
        int t = tile_;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int ox = tx * BW + xx, oy = ty * BH + yy, n = t * BNt + nn;
        if (ox < p.Wo && oy < p.Ho && n < dyn.N) {
          const size_t rpix = p.res_mode == RES_SAME ? ((size_t)n * p.Ho + oy) * p.Wo + ox
                                                     : ((size_t)n * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix * p.Cout + nb * BLOCK_N + g_ * gcols);
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = __ldg(rp + j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      int tile = blockIdx.x + (BLOCK_N == 64 ? half * (int)gridDim.x : 0);
      int it = BLOCK_N == 64 ? half : 0;                  // index of `tile` among this CTA's tiles
      if (use_res && p.res_tma && tile < total_tiles) res_issue(tile, g0);
      for (; tile < total_tiles; tile += step, it += (BLOCK_N == 64 ? 2 : 1)) {
        int t = tile;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int as = it % acc_n;
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
        uint4 rv[8];
        if (use_res && !p.res_tma) res_load(tile, g0, rv);
        TMR_WAIT(tmr_wait, mbar_wait(tfull0 + 8 * as, (uint32_t)((it / acc_n) & 1)))
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int g = g0; g < g1; ++g) {
          if (use_res && p.res_tma) {
            mbar_wait(rbar, rphase); rphase ^= 1;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[j].x), "=r"(rv[j].y), "=r"(rv[j].z), "=r"(rv[j].w)
                           : "r"(raddr + (((uint32_t)j ^ rsw) << 4)) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (g + 1 < g1) res_issue(tile, g + 1);
            else if (tile + step < total_tiles) res_issue(tile + step, g0);
          } else if (use_res && g > g0) {
            res_load(tile, g, rv);
          }
          epilogue_group_tma<BLOCK_N>(p, &maps.o, stg, pstg, tmem_acc, q, lane, g, use_res, rv, nb, tx * BW + (m0 & (BW - 1)),
                                      ty * BH + ((m0 >> p.lw) & (BH - 1)), t * BNt + (m0 >> (p.lw + p.lh)));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      }
    } else {
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile;
      const int nb = t % p.n_blocks; t /= p.n_blocks;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int ox = tx * BW + xx, oy = ty * BH + yy, n = t * BNt + nn;
      const bool valid = ox < p.Wo && oy < p.Ho && n < dyn.N;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
      const uint32_t tfull_bar = tfull0 + 8 * as, parity = aphase;
      if constexpr (MODE == MODE_DBHEAD) {
        mbar_wait(tfull_bar, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue_dbhead(p, head_s + nb * 512, ex.b2[nb], ex.thr, nb, tmem_acc, q, half, valid, n, oy, ox);
      } else if constexpr (MODE == MODE_LSTM) {
        lstm_prefetch(p, half, ox, nb);
        mbar_wait(tfull_bar, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue_lstm(p, tmem_acc, q, half, ox, nb);
      } else {
        const size_t opix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
        size_t rpix = 0;
        if (p.res_mode == RES_SAME) rpix = opix;
        else if (p.res_mode == RES_UP2) rpix = ((size_t)n * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
#ifdef VTD_TIMERS
        TMR_WAIT(tmr_wait, mbar_wait(tfull_bar, parity))
#endif
        epilogue_conv<BLOCK_N>(p, tmem_acc, q, half, valid, opix, rpix, nb, tfull_bar, parity);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == acc_n) { as = 0; aphase ^= 1; }
    }
    }
    if (p.epi_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    if (warp == 2) { TMR_STORE(2) }
  }

  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// ---- CTA-pair variant of the halo convolution (3x3 s1 p1, streamed weights, n_blocks == 1) -------------------------------
// The single-CTA N = 128 / N = 256 tiles are bound by shared-memory bandwidth: per SS-form MMA the tensor core reads A (4 KB)
// and B (N x 32 B) from shared memory while TMA writes the same slabs in (profiles/r01_ncu_notes.md, finding 9).  With
// tcgen05.mma.cta_group::2 a cluster of two CTAs computes TWO spatial tiles (one per CTA: its own patch, its own 128 TMEM
// lanes) against ONE copy of the weights split between them: CTA r streams and keeps rows r*N/2.. of every weight slab,
// the tensor cores of both SMs read both halves.  Weight bytes written and read per SM halve.  (profiles/micro/umma_2cta.cu
// checks the operand placement.)  Roles per CTA: warp 0 TMA producer (own patch + own weight half; every load completes
// on the LEADER's full barrier), warp 1 MMA issuer (leader only; commits are multicast to both CTAs' empty / tfull
// barriers), warps 2..9 epilogue (own TMEM; arrive on the leader's tempty).
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {          // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int HALF_N = BLOCK_N / 2;
  constexpr int BSLAB = HALF_N * 128;                    // one tap of this CTA's weight half
  constexpr int ACC = 512 / BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  const int a_st = p.stages, b_st = p.b_stages, taps = p.b_taps;
  const uint32_t bslot = (uint32_t)taps * BSLAB;
  const uint32_t bring = ring + a_st * HALO_SLOT;
  const uint32_t stg0 = bring + b_st * bslot;
  const uint32_t pstg0 = stg0 + NUM_EPI_WARPS * 4096;    // pooled boxes (p.pool only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + (pstg0 - ring) + (p.pool ? NUM_EPI_WARPS * 2048 : 0));
  const uint32_t afull0 = smem_u32(bars), aempty0 = afull0 + 8 * 4, bfull0 = aempty0 + 8 * 4, bempty0 = bfull0 + 8 * 8;
  const uint32_t tfull0 = bempty0 + 8 * 8, tempty0 = tfull0 + 8 * 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a_st; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    for (int i = 0; i < b_st; ++i) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
    for (int i = 0; i < ACC; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[1]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b4) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.o) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // both CTAs' barriers exist before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = p.Cin / BLOCK_K;
  const int tgroups = 9 / taps;
  const DynCount dyn = dyn_count(p);
  const int total_tiles = dyn.total_tiles;
  const int npairs = (total_tiles + 1) >> 1;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  auto my_tile = [&](int pair) { const int t = 2 * pair + (int)rank; return t < total_tiles ? t : total_tiles - 1; };   // odd tail: the peer repeats the last tile

  if (warp == 0) {
    // ===================== TMA producer =====================
    int as_ = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
    // p.bres (one 64-channel chunk, the ring holds all nine taps of this CTA's weight half): loaded once, never released --
    // streamed per tile the weights of a 64 -> 128 layer alone would ask more of L2 than it delivers
    if (p.bres && cluster_id < npairs && elect_one())
      for (int tg = 0; tg < tgroups; ++tg) {
        if (rank == 0) mbar_expect_tx(bfull0 + 8 * tg, 2u * bslot);
        tma_load_4d_2sm(bring + tg * bslot, &maps.b4, bfull0 + 8 * tg, 0, (int)rank * HALF_N, 0, tg * taps);
      }
    __syncwarp();
    for (int pair = cluster_id; pair < npairs; pair += nclusters) {
      int t = my_tile(pair);
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int x0 = tx * 8, y0 = ty * 16, n0 = t;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(aempty0 + 8 * as_, aph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(afull0 + 8 * as_, 2u * HALO_BYTES);          // both CTAs' patches complete on the leader's barrier
          tma_load_4d_2sm(ring + as_ * HALO_SLOT, &maps.a[1], afull0 + 8 * as_, kc * BLOCK_K, x0 - 1, y0 - 1, n0);
        }
        __syncwarp();
        if (++as_ == a_st) { as_ = 0; aph ^= 1; }
        if (p.bres) continue;
        for (int tg = 0; tg < tgroups; ++tg) {
          mbar_wait(bempty0 + 8 * bs, bph ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bfull0 + 8 * bs, 2u * bslot);
            tma_load_4d_2sm(bring + bs * bslot, &maps.b4, bfull0 + 8 * bs, 0, (int)rank * HALF_N, kc, tg * taps);
          }
          __syncwarp();
          if (++bs == b_st) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (VTD_UMMA_AB_FMT << 7) | (VTD_UMMA_AB_FMT << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int as_ = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int acc = 0; uint32_t accph = 0;
      for (int pair = cluster_id; pair < npairs; pair += nclusters) {
        mbar_wait(tempty0 + 8 * acc, accph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(afull0 + 8 * as_, aph);
          const uint32_t sa = ring + as_ * HALO_SLOT;
          int fr = 0, fs = 0;
          for (int tg = 0; tg < tgroups; ++tg) {
            if (!p.bres || pair == cluster_id) mbar_wait(bfull0 + 8 * bs, bph);       // resident weights: landed once
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const uint64_t ad0 = umma_desc_sw128(sa + (uint32_t)(fr * HALO_PW + fs) * 128u, HALO_PW * 128u);
              const uint64_t bd0 = umma_desc<128>(bring + bs * bslot);
              for (int tt = 0; tt < taps; ++tt)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_2sm(d_tmem, ad0 + (uint64_t)(tt * 8 + k * 2), bd0 + (uint64_t)(tt * (BSLAB >> 4) + k * 2), idesc,
                               (kc | tg | tt | k) ? 1u : 0u);
              if (!p.bres) umma_commit_2sm(bempty0 + 8 * bs);
              if (tg == tgroups - 1) {
                umma_commit_2sm(aempty0 + 8 * as_);
                if (kc == kchunks - 1) umma_commit_2sm(tfull0 + 8 * acc);
              }
            }
            __syncwarp();
            if (taps == 3) ++fr; else if (++fs == 3) { fs = 0; ++fr; }
            if (++bs == b_st) { bs = 0; bph ^= 1; }
          }
          if (++as_ == a_st) { as_ = 0; aph ^= 1; }
        }
        if (++acc == ACC) { acc = 0; accph ^= 1; }
      }
    }
  } else if (warp < 2 + NUM_EPI_WARPS) {
    // ===================== epilogue =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int xx = m & 7, yy = (m >> 3) & 15;
    const int groups = BLOCK_N * 2 / 128;
    const int g0 = half * (groups / 2), g1 = g0 + groups / 2;
    const bool use_res = p.res_mode != RES_NONE;
    const uint32_t stg = stg0 + (uint32_t)(warp - 2) * 4096u;
    const int m0 = q * 32;
    uint32_t tempty_leader;                              // the leader's tempty barrier, cluster address (own when rank == 0)
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(tempty_leader) : "r"(tempty0), "r"(0u));
    int acc = 0; uint32_t accph = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters) {
      int t = my_tile(pair);
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int ox = tx * 8 + xx, oy = ty * 16 + yy, n = t;
      const bool valid = ox < p.Wo && oy < p.Ho && n < dyn.N;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N);
      mbar_wait(tfull0 + 8 * acc, accph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int g = g0; g < g1; ++g) {
        uint4 rv[8];
        if (use_res) {
          if (valid) {
            const size_t rpix = p.res_mode == RES_SAME ? ((size_t)n * p.Ho + oy) * p.Wo + ox
                                                       : ((size_t)n * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix * p.Cout + g * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = __ldg(rp + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        epilogue_group_tma<BLOCK_N>(p, &maps.o, stg, pstg0 + (uint32_t)(warp - 2) * 2048u, tmem_acc, q, lane, g, use_res, rv, 0,
                                    tx * 8 + (m0 & 7), ty * 16 + ((m0 >> 3) & 15), t);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader + 8 * acc) : "memory");
      if (++acc == ACC) { acc = 0; accph ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // nobody exits while the peer may still address its memory
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- CTA-pair variant of the GENERAL implicit-GEMM convolution (N = 256 tiles, streamed weights) -----------------------------
// Same idea as conv_tc2_kernel for the layers that cannot run in halo mode (CRNN 8x32 / 4x32 maps, stride-2 and 1x1 layers,
// the LSTM input projections): per K step every CTA loads the tap-shifted box of ITS tile (16 KB) and ITS half of the
// weight slab (128 rows, 16 KB); one tcgen05.mma.cta_group::2 stream of the leader multiplies both tiles.  Tiles of a pair
// share the N block.
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc2g_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int BLOCK_N = 256, HALF_N = 128;
  constexpr int A_BYTES = BLOCK_M * 128, BH_BYTES = HALF_N * 128, SLOT = A_BYTES + BH_BYTES;
  constexpr int ACC = 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  const int stages = p.stages;
  const uint32_t stg0 = ring + stages * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + (stg0 - ring) + NUM_EPI_WARPS * 4096 + (p.pool ? NUM_EPI_WARPS * 2048 : 0));
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * 8, tfull0 = empty0 + 8 * 8, tempty0 = tfull0 + 8 * 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < ACC; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b4) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.o) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int BW = 1 << p.lw, BH = 1 << p.lh;
  const int BNt = BLOCK_M >> (p.lw + p.lh);
  const int kchunks = p.Cin / BLOCK_K;
  const int ksteps = p.KH * p.KW * kchunks;
  const DynCount dyn = dyn_count(p);
  const int spatial = p.tiles_x * p.tiles_y * dyn.tiles_n;
  const int spairs = (spatial + 1) >> 1;
  const int npairs = spairs * p.n_blocks;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  // pair -> (N block, this CTA's spatial tile); the peer of an odd tail repeats the last tile
  auto decode = [&](int pair, int& nb, int& tx, int& ty, int& tn) {
    nb = pair % p.n_blocks;
    int t = 2 * (pair / p.n_blocks) + (int)rank;
    if (t >= spatial) t = spatial - 1;
    tx = t % p.tiles_x; t /= p.tiles_x;
    ty = t % p.tiles_y; tn = t / p.tiles_y;
  };

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters) {
      int nb, tx, ty, tn;
      decode(pair, nb, tx, ty, tn);
      const int x0 = tx * BW, y0 = ty * BH, n0 = tn * BNt;
      for (int r = 0; r < p.KH; ++r)
        for (int sx = 0; sx < p.KW; ++sx) {
          int mi = 0, xo, yo;
          if (p.stride == 1) { xo = sx - p.pad; yo = r - p.pad; }
          else {
            const int tyy = r - p.pad, txx = sx - p.pad;
            const int py = tyy & 1, px = txx & 1;
            mi = py * 2 + px; yo = (tyy - py) / 2; xo = (txx - px) / 2;
          }
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            if (elect_one()) {
              const uint32_t sa = ring + stage * SLOT, fb = full0 + 8 * stage;
              if (rank == 0) mbar_expect_tx(fb, 2u * SLOT);
              tma_load_4d_2sm(sa, &maps.a[mi], fb, kc * BLOCK_K, x0 + xo, y0 + yo, n0);
              tma_load_2d_2sm(sa + A_BYTES, &maps.b4, fb, (r * p.KW + sx) * p.Cin + kc * BLOCK_K, nb * BLOCK_N + (int)rank * HALF_N);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (VTD_UMMA_AB_FMT << 7) | (VTD_UMMA_AB_FMT << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t accph = 0;
      bool ready = false;
      for (int pair = cluster_id; pair < npairs; pair += nclusters) {
        mbar_wait(tempty0 + 8 * acc, accph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int ks = 0; ks < ksteps; ++ks) {
          if (!ready) mbar_wait(full0 + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int nstage = stage + 1 == stages ? 0 : stage + 1;
          const uint32_t nphase = stage + 1 == stages ? phase ^ 1 : phase;
          const bool probe = mbar_test(full0 + 8 * nstage, nphase);
          if (elect_one()) {
            const uint32_t sa = ring + stage * SLOT;
            const uint64_t ad = umma_desc<128>(sa), bd = umma_desc<128>(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_2sm(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ks | k) ? 1u : 0u);
            umma_commit_2sm(empty0 + 8 * stage);
            if (ks == ksteps - 1) umma_commit_2sm(tfull0 + 8 * acc);
          }
          ready = __any_sync(0xffffffffu, probe);
          stage = nstage; phase = nphase;
        }
        if (++acc == ACC) { acc = 0; accph ^= 1; }
      }
    }
  } else if (warp < 2 + NUM_EPI_WARPS) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int xx = m & (BW - 1), yy = (m >> p.lw) & (BH - 1), nn = m >> (p.lw + p.lh);
    const int groups = (p.out_f32 ? BLOCK_N * 4 : BLOCK_N * 2) / 128;
    const int gcols = p.out_f32 ? 32 : 64;
    const int g0 = half * (groups / 2), g1 = g0 + groups / 2;
    const bool use_res = p.res_mode != RES_NONE && !p.out_f32;
    const uint32_t stg = stg0 + (uint32_t)(warp - 2) * 4096u;
    const int m0 = q * 32;
    uint32_t tempty_leader;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(tempty_leader) : "r"(tempty0), "r"(0u));
    int acc = 0; uint32_t accph = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters) {
      int nb, tx, ty, tn;
      decode(pair, nb, tx, ty, tn);
      const int ox = tx * BW + xx, oy = ty * BH + yy, n = tn * BNt + nn;
      const bool valid = ox < p.Wo && oy < p.Ho && n < dyn.N;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N);
      mbar_wait(tfull0 + 8 * acc, accph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int g = g0; g < g1; ++g) {
        uint4 rv[8];
        if (use_res) {
          if (valid) {
            const size_t rpix = p.res_mode == RES_SAME ? ((size_t)n * p.Ho + oy) * p.Wo + ox
                                                       : ((size_t)n * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix * p.Cout + nb * BLOCK_N + g * gcols);
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = __ldg(rp + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        epilogue_group_tma<BLOCK_N>(p, &maps.o, stg, stg0 + (uint32_t)NUM_EPI_WARPS * 4096u + (uint32_t)(warp - 2) * 2048u, tmem_acc, q,
                                    lane, g, use_res, rv, nb, tx * BW + (m0 & (BW - 1)), ty * BH + ((m0 >> p.lw) & (BH - 1)),
                                    tn * BNt + (m0 >> (p.lw + p.lh)));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_leader + 8 * acc) : "memory");
      if (++acc == ACC) { acc = 0; accph ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- one-pass DB head (text_detector.py:58-86 + `> threshold`, :144) -----------------------------------------------------------
// The whole DBHead of both branches in ONE kernel and one HBM pass: P2 is read once, probability, threshold and mask are
// written once, and the 128-channel feature map between the 3x3 convolutions and the transposed convolutions never
// exists in global memory.  Built on the CTA-pair halo convolution above (conv_tc2_kernel<128>: the two branches' 3x3
// 256->64 convolutions merged into one 256->128 implicit GEMM, BN folded, ReLU):
//   1. main loop as conv_tc2_kernel<128>: accumulator [128 px x 128 ch] per CTA in TMEM (two buffers, columns 0..255);
//   2. the 16 epilogue warps add the bias, apply ReLU, round to the 16-bit storage type and write the tile to shared
//      memory as the K-major SWIZZLE_128B A operand of a second GEMM (one 16 KB slab per branch) -- this IS the feature
//      map the unfused path stores, bit for bit;
//   3. ConvTranspose2d(64->64, k2, s2) + BN of one branch is the GEMM [px x 64] x [64 x 256] (256 = 2x2 output
//      positions x 64 channels): four more tcgen05.mma (cta_group::2, N = 256) into TMEM columns 256..511, issued by
//      the same MMA warp between the groups of the NEXT tile's convolution MMAs (the tail of tile i overlaps the
//      convolution of tile i+1; every wait of that warp keeps serving the tail, so the two pipelines cannot deadlock);
//      W1 (this CTA's 128 rows of each branch, 32 KB) stays resident in shared memory;
//   4. the epilogue warps read that accumulator: + folded BN shift, ReLU, the second transposed convolution (64 -> 1,
//      k2, s2: four 64-long dot products per position, fp32 FFMA2, weights in shared memory), + optional logit plane,
//      sigmoid, `> thr`.  Warp (quarter q, position g = dy*2+dx) owns 32 pixels and one of the four positions: it writes
//      the 2x2 block of each pixel's 4x4 output block.  The same arithmetic in the same order as MODE_DBHEAD, so the maps
//      are bit-identical to the unfused path's.
// Roles per CTA (18 warps): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..17 epilogue.
constexpr int HF_EPI_WARPS = 16;
constexpr int HF_THREADS = 32 * (2 + HF_EPI_WARPS);
constexpr int HF_A2_BYTES = 2 * 128 * 128;        // ReLU(feat) of the tile: 2 branches x 128 px x 64 ch x 2 B
constexpr int HF_W1_BYTES = 2 * 128 * 128;        // this CTA's half of W1: 2 branches x 128 rows x 64 K x 2 B
constexpr int HF_CONST_BYTES = 4096;              // b1 [2][256] + w2 [2][64][4] fp32

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HF_THREADS, 1)
dbhead_fused_kernel(const __grid_constant__ TcMaps maps, const TcParams p, const __grid_constant__ HeadConsts ex) {
  constexpr int BLOCK_N = 128, HALF_N = 64;
  constexpr int BSLAB = HALF_N * 128;                    // one tap of this CTA's weight half
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  const int a_st = p.stages, b_st = p.b_stages, taps = p.b_taps;
  const uint32_t bslot = (uint32_t)taps * BSLAB;
  const uint32_t bring = ring + a_st * HALO_SLOT;
  const uint32_t a2_0 = bring + b_st * bslot;            // 1024-aligned: HALO_SLOT and bslot are multiples of 1024
  const uint32_t w1_0 = a2_0 + HF_A2_BYTES;
  const uint32_t const_0 = w1_0 + HF_W1_BYTES;
  float* head_s = reinterpret_cast<float*>(ring_ptr + (const_0 - ring));                 // [2][256 b1 + 256 w2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + (const_0 - ring) + HF_CONST_BYTES);
  const uint32_t afull0 = smem_u32(bars), aempty0 = afull0 + 8 * 4, bfull0 = aempty0 + 8 * 4, bempty0 = bfull0 + 8 * 8;
  const uint32_t tfull0 = bempty0 + 8 * 8, tempty0 = tfull0 + 8 * 2;
  const uint32_t a2full = tempty0 + 8 * 2, a2empty = a2full + 8, d2full = a2empty + 8, d2empty = d2full + 8, w1full = d2empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  for (int i = threadIdx.x; i < 2 * 512; i += HF_THREADS) {
    const int hd = i >> 9, r = i & 511;
    head_s[i] = r < 256 ? ex.b1[hd][r] : (&ex.w2[hd][0][0])[r - 256];
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a_st; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    for (int i = 0; i < b_st; ++i) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * HF_EPI_WARPS); }
    mbar_init(a2full, 2 * HF_EPI_WARPS); mbar_init(a2empty, 1);
    mbar_init(d2full, 1); mbar_init(d2empty, 2 * HF_EPI_WARPS);
    mbar_init(w1full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[1]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b4) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // both CTAs' barriers exist before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = p.Cin / BLOCK_K;
  const int tgroups = 9 / taps;
  const int total_tiles = p.total_tiles;
  const int npairs = (total_tiles + 1) >> 1;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int my_pairs = cluster_id < npairs ? (npairs - cluster_id + nclusters - 1) / nclusters : 0;   // tiles this CTA computes
  auto my_tile = [&](int pair) { const int t = 2 * pair + (int)rank; return t < total_tiles ? t : total_tiles - 1; };   // odd tail: the peer repeats the last tile

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (my_pairs > 0) {
      if (elect_one()) {
        // W1: rows rank*128.. of each branch's [256 x 64] matrix; both CTAs' loads complete on the leader's barrier
        if (rank == 0) mbar_expect_tx(w1full, 2u * HF_W1_BYTES);
        tma_load_2d_2sm(w1_0, &maps.b, w1full, 0, (int)rank * 128);
        tma_load_2d_2sm(w1_0 + 128 * 128, &maps.b, w1full, 0, 256 + (int)rank * 128);
      }
      __syncwarp();
    }
    int as_ = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters) {
      int t = my_tile(pair);
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int x0 = tx * 8, y0 = ty * 16, n0 = t;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(aempty0 + 8 * as_, aph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(afull0 + 8 * as_, 2u * HALO_BYTES);
          tma_load_4d_2sm(ring + as_ * HALO_SLOT, &maps.a[1], afull0 + 8 * as_, kc * BLOCK_K, x0 - 1, y0 - 1, n0);
        }
        __syncwarp();
        if (++as_ == a_st) { as_ = 0; aph ^= 1; }
        for (int tg = 0; tg < tgroups; ++tg) {
          mbar_wait(bempty0 + 8 * bs, bph ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bfull0 + 8 * bs, 2u * bslot);
            tma_load_4d_2sm(bring + bs * bslot, &maps.b4, bfull0 + 8 * bs, 0, (int)rank * HALF_N, kc, tg * taps);
          }
          __syncwarp();
          if (++bs == b_st) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && my_pairs > 0) {
      const uint32_t idesc = (1u << 4) | (VTD_UMMA_AB_FMT << 7) | (VTD_UMMA_AB_FMT << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (VTD_UMMA_AB_FMT << 7) | (VTD_UMMA_AB_FMT << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t d2_tmem = tmem_base + 256u;
      // tail pipeline: use u = 2 * tile + branch of the D2 accumulator.  MMA2(u) may be issued once the tile's A operand
      // is in shared memory (branch 0: a2full) and the epilogue has drained the previous use of D2 (d2empty).
      const int n_uses = 2 * my_pairs;
      int u = 0;
      bool w1_ready = false;
      auto serve_tail = [&]() -> bool {                          // non-blocking; whole warp converged; true if it issued
        if (u >= n_uses) return false;
        const int tile = u >> 1, br = u & 1;
        bool ok = mbar_test(d2empty, (uint32_t)((u & 1) ^ 1));
        if (br == 0) ok = ok && mbar_test(a2full, (uint32_t)(tile & 1));
        if (!w1_ready) ok = ok && mbar_test(w1full, 0);
        if (!__all_sync(0xffffffffu, ok)) return false;
        w1_ready = true;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t ad = umma_desc<128>(a2_0 + (uint32_t)br * (128u * 128u));
          const uint64_t bd = umma_desc<128>(w1_0 + (uint32_t)br * (128u * 128u));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_2sm(d2_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc2, k ? 1u : 0u);
          umma_commit_2sm(d2full);
          if (br == 1) umma_commit_2sm(a2empty);                 // both branches have read the tile's A operand
        }
        __syncwarp();
        ++u;
        return true;
      };
      auto wait_serving = [&](uint32_t bar, uint32_t parity) {   // a blocking wait that keeps the tail pipeline moving
        uint32_t spins = 0;
        while (!__any_sync(0xffffffffu, mbar_test(bar, parity))) {
          if (!serve_tail() && ++spins == (1u << 28)) __trap();
        }
      };
      int as_ = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int acc = 0; uint32_t accph = 0;
      // p.kps (here: groups in flight, 0 = unlimited): the tensor pipe executes in issue order, so a tail MMA waits behind
      // every convolution group already issued; bounding the run-ahead bounds that wait (the weight ring still prefetches)
      const int max_ahead = p.kps;
      int gslot = 0; uint32_t gph = 0; int gcount = 0;          // slot / phase of the group issued max_ahead groups ago
      for (int pair = cluster_id; pair < npairs; pair += nclusters) {
        wait_serving(tempty0 + 8 * acc, accph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kc = 0; kc < kchunks; ++kc) {
          wait_serving(afull0 + 8 * as_, aph);
          const uint32_t sa = ring + as_ * HALO_SLOT;
          int fr = 0, fs = 0;
          for (int tg = 0; tg < tgroups; ++tg) {
            wait_serving(bfull0 + 8 * bs, bph);
            if (max_ahead > 0) {
              if (gcount >= max_ahead) {                           // the group issued max_ahead groups ago has left the tensor pipe
                wait_serving(bempty0 + 8 * gslot, gph);
                if (++gslot == b_st) { gslot = 0; gph ^= 1; }
              } else ++gcount;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const uint64_t ad0 = umma_desc_sw128(sa + (uint32_t)(fr * HALO_PW + fs) * 128u, HALO_PW * 128u);
              const uint64_t bd0 = umma_desc<128>(bring + bs * bslot);
              for (int tt = 0; tt < taps; ++tt)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_2sm(d_tmem, ad0 + (uint64_t)(tt * 8 + k * 2), bd0 + (uint64_t)(tt * (BSLAB >> 4) + k * 2), idesc,
                               (kc | tg | tt | k) ? 1u : 0u);
              umma_commit_2sm(bempty0 + 8 * bs);
              if (tg == tgroups - 1) {
                umma_commit_2sm(aempty0 + 8 * as_);
                if (kc == kchunks - 1) umma_commit_2sm(tfull0 + 8 * acc);
              }
            }
            __syncwarp();
            serve_tail();                                          // the previous tile's transposed convolutions, in between
            if (taps == 3) ++fr; else if (++fs == 3) { fs = 0; ++fr; }
            if (++bs == b_st) { bs = 0; bph ^= 1; }
          }
          if (++as_ == a_st) { as_ = 0; aph ^= 1; }
        }
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
      uint32_t spins = 0;
      while (u < n_uses) { if (!serve_tail() && ++spins == (1u << 28)) __trap(); }      // drain
    }
  } else {
    // ===================== epilogue (16 warps) =====================
    const int q = warp & 3;                               // TMEM lane quarter
    const int sub = (warp - 2) >> 2;                      // 0..3: 32-column slice of the conv tile / output position g
    const int m = q * 32 + lane;
    const int xx = m & 7, yy = (m >> 3) & 15;
    uint32_t tempty_l, a2full_l, d2empty_l;               // the leader's barriers, cluster addresses (own when rank == 0)
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(tempty_l) : "r"(tempty0), "r"(0u));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a2full_l) : "r"(a2full), "r"(0u));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d2empty_l) : "r"(d2empty), "r"(0u));
    const int Wd = 4 * p.Wo, Hd = 4 * p.Ho;
    const int dy = sub >> 1, dx = sub & 1;
    int it = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters, ++it) {
      int t = my_tile(pair);
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; t /= p.tiles_y;
      const int ox = tx * 8 + xx, oy = ty * 16 + yy, n = t;
      const bool valid = ox < p.Wo && oy < p.Ho && n < p.N;
      const int acc = it & 1;
      // ---- step 1: conv accumulator -> bias, ReLU, 16-bit -> A operand of the transposed-convolution GEMM
      mbar_wait(tfull0 + 8 * acc, (uint32_t)((it >> 1) & 1));
      mbar_wait(a2empty, (uint32_t)((it & 1) ^ 1));       // the previous tile's MMA2s have read the slabs
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + sub * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float* bias = p.bias + sub * 32;
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * j)), b1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * j + 4));
          bf16x2 h0 = pack2(fmaxf(__uint_as_float(v[8 * j]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.f));
          bf16x2 h1 = pack2(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.f));
          bf16x2 h2 = pack2(fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.f));
          bf16x2 h3 = pack2(fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.f));
          o[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                            *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
        // slab = branch (sub >> 1); row = pixel m (128 bytes = the branch's 64 channels); 16-byte chunk c of the row at c ^ (m & 7)
        const uint32_t row = a2_0 + (uint32_t)(sub >> 1) * (128u * 128u) + (uint32_t)m * 128u;
        const uint32_t sw = (uint32_t)(m & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c = (uint32_t)((sub & 1) * 4 + j);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ sw) << 4)), "r"(o[j].x), "r"(o[j].y),
                       "r"(o[j].z), "r"(o[j].w) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core's reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        // a2full publishes shared-memory writes to the pair's tensor cores: release.  tempty / d2empty only say "my
        // tcgen05.ld have completed" (tcgen05.wait::ld + fence above): relaxed -- a release here would also wait for
        // every global store of the previous branch to be performed (19 % of the kernel's stall samples, ncu).
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a2full_l) : "memory");
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tempty_l + 8 * acc) : "memory");
      }
      // ---- steps 2, 3: per branch, ConvT1 accumulator -> +shift, ReLU -> ConvT2 -> (+logit plane) -> sigmoid -> maps
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
        float2 lbv[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (br == 0 && p.logit_bias && valid) {            // requested before the wait: the L2 / DRAM latency hides behind the MMAs
#pragma unroll
          for (int dy2 = 0; dy2 < 2; ++dy2)
            lbv[dy2] = __ldg(reinterpret_cast<const float2*>(p.logit_bias + ((size_t)n * Hd + 4 * oy + 2 * dy + dy2) * Wd + 4 * ox + 2 * dx));
        }
        mbar_wait(d2full, (uint32_t)br);                   // use u = 2 * it + br: parity u & 1 = br
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* hs = head_s + br * 512;
        float o4[4];
        const float b2 = ex.b2[br];
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = b2;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(sub * 64 + ch * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* b1 = hs + sub * 64 + ch * 32;
          const float4* w2 = reinterpret_cast<const float4*>(hs + 256) + ch * 32;
          float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 h = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                  *reinterpret_cast<const float2*>(b1 + j));
            h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f);
            const float4 w0 = w2[j], w1 = w2[j + 1];
            a01 = __ffma2_rn(make_float2(h.x, h.x), make_float2(w0.x, w0.y), a01);
            a23 = __ffma2_rn(make_float2(h.x, h.x), make_float2(w0.z, w0.w), a23);
            a01 = __ffma2_rn(make_float2(h.y, h.y), make_float2(w1.x, w1.y), a01);
            a23 = __ffma2_rn(make_float2(h.y, h.y), make_float2(w1.z, w1.w), a23);
          }
          o4[0] += a01.x; o4[1] += a01.y; o4[2] += a23.x; o4[3] += a23.y;
        }
        // D2 has been read: the next MMA2 may overwrite it while the stores below drain
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(d2empty_l) : "memory");
        if (valid) {
          float* __restrict__ outp = br == 0 ? p.prob : p.thresh;
#pragma unroll
          for (int dy2 = 0; dy2 < 2; ++dy2) {
            float v0 = o4[dy2 * 2 + 0], v1 = o4[dy2 * 2 + 1];
            const size_t oidx = ((size_t)n * Hd + 4 * oy + 2 * dy + dy2) * Wd + 4 * ox + 2 * dx;
            if (br == 0 && p.logit_bias) {
              const float2 lb = lbv[dy2];
              v0 += lb.x; v1 += lb.y;
            }
            v0 = 1.0f / (1.0f + expf(-v0)); v1 = 1.0f / (1.0f + expf(-v1));
            *reinterpret_cast<float2*>(outp + oidx) = make_float2(v0, v1);
            if (br == 0)
              *reinterpret_cast<uint16_t*>(p.mask + oidx) = (uint16_t)((v0 > ex.thr ? 1u : 0u) | (v1 > ex.thr ? 0x100u : 0u));
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // nobody exits while the peer may still address its memory
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- DBNet stem + max-pool in one kernel ---------------------------------------------------------------------------------------
// conv1 7x7 s2 (+ folded BN + ReLU) and nn.MaxPool2d(3, 2, 1) (torchvision ResNet children[0..3], text_detector.py:17-19) without
// the full-resolution stem map ever reaching HBM: the separate kernels wrote 494 MB and re-read them per 16 frames at
// 736x1312; this one reads the padded input (125 MB) and writes the pooled map (124 MB).
// A work item is (image, column tile, band of pooled rows).  Stem rows are computed one after the other exactly as the
// direct-window stem does (MODE_WIN / win2 above: the 7 padded input rows under a row tile of 128 outputs are bulk-copied
// once, the MMA reads the overlapping 64-byte windows in place through a no-swizzle descriptor; 14 MMAs 128x64x16 per row
// tile), but the epilogue keeps the last two ReLU'd rows of its column in REGISTERS and, after every odd stem row 2y+1,
// emits pooled row y = max over stem rows 2y-1..2y+1 (registers) and stem columns 2x-1..2x+1 (row maxima staged once in
// shared memory).  A column tile of 128 stem columns
// starts at stem column 126 t - 1 and owns 63 pooled columns, so every pooling window lies inside its tile (2 of 128
// columns are computed twice); out-of-range rows / columns contribute zeros, which is neutral after ReLU.
// Roles: warp 0 producer (bulk copies), warp 1 MMA issuer, warps 2..9 epilogue + pooling.
constexpr int SP_EPI_WARPS = 8;
constexpr int SP_THREADS = 32 * (2 + SP_EPI_WARPS);
constexpr int SP_ROWBUF = 128 * 128;                 // one stem row tile: 128 px x 64 ch x 2 B
// The same kernel also runs the CRNN's first layer (text_recognizer.py:17: Conv2d(3, 64, 3, 1, 1) + BN + ReLU + MaxPool2d(2, 2)),
// template parameter CRNN: 3 filter rows instead of 7, conv stride 1 (8 channels per padded pixel, so consecutive outputs are
// again 16 bytes apart), 2x2 s2 pooling without overlap (column tiles of 128 conv columns = 64 pooled ones).

struct StemPoolParams {
  const uint8_t* in;          // padded input [N][Hin][Win][4] 16-bit: 3 rows above, 6 px left of the image
  long long in_rp, in_ip;     // row / image pitch in bytes
  const float* bias;
  bf16* out;                  // pooled map [N][Hp][Wp][64]
  int N, Ho, Wo, Hp, Wp;      // stem map Ho x Wo, pooled map Hp x Wp
  int tiles_x, bands, rows_per_band, total_items, stages;
  const int* n_dyn; int n_first;   // CRNN first layer: crop count on the device (see TcParams::n_dyn)
};

template <bool CRNN>
__global__ void __launch_bounds__(SP_THREADS, 1) stem_pool_kernel(const __grid_constant__ CUtensorMap wmap, const StemPoolParams p) {
  constexpr int B_ROW_BYTES = 64 * 64;                 // weights of one filter row: 64 out channels x 32 K (64-byte rows, SWIZZLE_64B)
  constexpr int NR = CRNN ? 3 : 7;                     // filter rows = bulk-copied input rows per conv row tile
  constexpr int SY = CRNN ? 1 : 2;                     // input rows per conv row (conv stride)
  constexpr int SP_POOLED = CRNN ? 64 : 63;            // pooled columns per column tile
  constexpr int SLOT = CRNN ? 7 * 1024 : WIN2_SLOT;    // ring slot: NR slabs of WIN2_SLAB bytes
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  const int stages = p.stages;
  const uint32_t w0 = ring + stages * SLOT;            // resident weights, NR x 4 KB
  const uint32_t rows0 = w0 + NR * B_ROW_BYTES;        // two buffers for the row maxima (DBNet's horizontal pooling)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + (rows0 - ring) + 2 * SP_ROWBUF);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * 16, tfull0 = empty0 + 8 * 16, tempty0 = tfull0 + 8 * 4, wfull = tempty0 + 8 * 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, SP_EPI_WARPS); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  int total_items = p.total_items;
  if (p.n_dyn) {
    int c = *p.n_dyn - p.n_first;
    c = c < 0 ? 0 : (c > p.N ? p.N : c);
    total_items = c * p.tiles_x * p.bands;
  }

  // item -> image, column tile, band; the stem rows it computes: first = max(2 ya - 1, 0) .. last = 2 yb - 1
  auto decode = [&](int item, int& n, int& t, int& ya, int& yb) {
    const int band = item % p.bands; item /= p.bands;
    t = item % p.tiles_x; n = item / p.tiles_x;
    ya = band * p.rows_per_band; yb = min(ya + p.rows_per_band, p.Hp);
  };

  if (warp == 0) {
    // ===================== producer =====================
    if ((int)blockIdx.x < total_items && elect_one()) {
      mbar_expect_tx(wfull, (uint32_t)NR * B_ROW_BYTES);
      for (int j = 0; j < NR; ++j) tma_load_2d(w0 + j * B_ROW_BYTES, &wmap, wfull, j * 32, 0);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int n, t, ya, yb; decode(item, n, t, ya, yb);
      const int x0 = CRNN ? 128 * t : SP_POOLED * 2 * t - 1;                      // first conv column of the tile
      const uint8_t* base = p.in + (size_t)n * p.in_ip + (CRNN ? (size_t)x0 * 16 : (size_t)(2 * x0 + 2) * 8);  // its first window byte in a padded row
      for (int r = CRNN ? 2 * ya : max(2 * ya - 1, 0); r <= 2 * yb - 1; ++r) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = ring + stage * SLOT, fb = full0 + 8 * stage;
          mbar_expect_tx(fb, (uint32_t)NR * WIN2_SLAB);
          const uint8_t* src = base + (size_t)(SY * r) * p.in_rp;
#pragma unroll
          for (int j = 0; j < NR; ++j)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sa + j * WIN2_SLAB), "l"(src + (size_t)j * p.in_rp), "r"((uint32_t)WIN2_SLAB), "r"(fb) : "memory");
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(64);
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    if ((int)blockIdx.x < total_items) mbar_wait(wfull, 0);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int n, t, ya, yb; decode(item, n, t, ya, yb);
      for (int r = CRNN ? 2 * ya : max(2 * ya - 1, 0); r <= 2 * yb - 1; ++r) {
        mbar_wait(tempty0 + 8 * as, aphase ^ 1);
        mbar_wait(full0 + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = ring + stage * SLOT;
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 64);
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const uint64_t ad = umma_desc_nosw(sa + j * WIN2_SLAB, 16, 128);
            const uint64_t bd = umma_desc<64>(w0 + j * B_ROW_BYTES);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_f16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (j | k) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * stage);
          umma_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
        if (++as == 4) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue + pooling (8 warps, 256 threads) =====================
    // A thread owns one conv column of the tile (TMEM lane) and 32 of the 64 channels in EVERY row, so the vertical
    // half of the pooling never leaves its registers: it keeps the previous two ReLU'd rows (packed 16-bit) and, after
    // every odd conv row, takes the maximum over the window's rows.  Only the horizontal half needs other threads'
    // columns: DBNet (3-wide windows that straddle warps) stages the row maxima in shared memory (two slots, one
    // CTA-wide barrier per POOLED row); the CRNN's 2-wide windows are lane pairs of one warp (one shuffle).
    const int q = warp & 3, half = (warp - 2) >> 2;        // TMEM lane quarter; which 32 of the 64 channels
    const int m = q * 32 + lane;                           // conv column inside the tile
    const int et = threadIdx.x - 64;                       // 0..255
    int as = 0; uint32_t aphase = 0;
    int slot = 0;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + half * 32 + j);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int n, t, ya, yb; decode(item, n, t, ya, yb);
      const int x0 = CRNN ? 128 * t : SP_POOLED * 2 * t - 1;
      const int col = x0 + m;
      const bool col_ok = col >= 0 && col < p.Wo;
      const int r_first = CRNN ? 2 * ya : max(2 * ya - 1, 0);
      uint32_t prev1[16], prev2[16];                        // rows r-1, r-2 of my column (zeros: neutral after ReLU)
#pragma unroll
      for (int j = 0; j < 16; ++j) { prev1[j] = 0u; prev2[j] = 0u; }
      // the accumulator of row r+1 is requested (tcgen05.ld) as soon as row r's has been turned into `cur`: its TMEM latency
      // hides behind the pooling work below
      uint32_t v[32];
      auto request = [&]() {
        mbar_wait(tfull0 + 8 * as, aphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 64 + half * 32), v);
      };
      request();
      for (int r = r_first; r <= 2 * yb - 1; ++r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
        if (++as == 4) { as = 0; aphase ^= 1; }
        uint32_t cur[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = col_ok ? fmaxf(__uint_as_float(v[2 * j]) + bias[2 * j], 0.f) : 0.f;
          const float b = col_ok ? fmaxf(__uint_as_float(v[2 * j + 1]) + bias[2 * j + 1], 0.f) : 0.f;
          bf16x2 h = pack2(a, b);
          cur[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (r < 2 * yb - 1) request();
        if ((r & 1) && r > 2 * ya - (CRNN ? 1 : 0)) {        // (DBNet: the band's warm-up row 2 ya - 1 is odd too and belongs to the band above)
          const int y = (r - 1) >> 1;                          // pooled row
          uint32_t vm[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            bf16x2 a = __hmax2(*reinterpret_cast<bf16x2*>(&cur[j]), *reinterpret_cast<bf16x2*>(&prev1[j]));
            if (!CRNN) a = __hmax2(a, *reinterpret_cast<bf16x2*>(&prev2[j]));
            vm[j] = *reinterpret_cast<uint32_t*>(&a);
          }
          if (CRNN) {
            // 2x2: columns 2x, 2x+1 are lanes 2i, 2i+1 of this warp; the even lane stores channels 0..15 of the pair's
            // maximum, the odd lane channels 16..31
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t o = __shfl_xor_sync(0xffffffffu, vm[j], 1);
              bf16x2 a = __hmax2(*reinterpret_cast<bf16x2*>(&vm[j]), *reinterpret_cast<const bf16x2*>(&o));
              vm[j] = *reinterpret_cast<uint32_t*>(&a);
            }
            const int xg = (x0 + m) >> 1;
            if (col_ok && xg < p.Wp) {
              uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)n * p.Hp + y) * p.Wp + xg) * 64 + half * 32 + (lane & 1) * 16);
              if (lane & 1) { dst[0] = make_uint4(vm[8], vm[9], vm[10], vm[11]); dst[1] = make_uint4(vm[12], vm[13], vm[14], vm[15]); }
              else { dst[0] = make_uint4(vm[0], vm[1], vm[2], vm[3]); dst[1] = make_uint4(vm[4], vm[5], vm[6], vm[7]); }
            }
          } else {
            const uint32_t rowb = rows0 + (uint32_t)slot * SP_ROWBUF;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t c = (uint32_t)(half * 4 + j);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (uint32_t)m * 128u + ((c ^ (uint32_t)(m & 7)) << 4)), "r"(vm[4 * j]),
                           "r"(vm[4 * j + 1]), "r"(vm[4 * j + 2]), "r"(vm[4 * j + 3]) : "memory");
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the row maxima of all 128 columns are in place (and every thread
                                                               // has left the pooling pass of the row before last: two slots)
            for (int idx = et; idx < SP_POOLED * 8; idx += 256) {
              const int xl = idx >> 3, c = idx & 7;
              const int xg = SP_POOLED * t + xl;
              if (xg >= p.Wp) continue;
              uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint32_t px = (uint32_t)(2 * xl + dx);
                uint4 tv;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(tv.x), "=r"(tv.y), "=r"(tv.z), "=r"(tv.w)
                             : "r"(rowb + px * 128u + (((uint32_t)c ^ (px & 7u)) << 4)) : "memory");
                bf16x2 a, b;
#define VTD_HMAX2(dst, src) a = *reinterpret_cast<bf16x2*>(&dst); b = *reinterpret_cast<bf16x2*>(&src); a = __hmax2(a, b); dst = *reinterpret_cast<uint32_t*>(&a);
                VTD_HMAX2(acc.x, tv.x) VTD_HMAX2(acc.y, tv.y) VTD_HMAX2(acc.z, tv.z) VTD_HMAX2(acc.w, tv.w)
#undef VTD_HMAX2
              }
              *reinterpret_cast<uint4*>(p.out + (((size_t)n * p.Hp + y) * p.Wp + xg) * 64 + c * 8) = acc;
            }
            slot ^= 1;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { prev2[j] = prev1[j]; prev1[j] = cur[j]; }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------

// tile shape: BN x BH x BW = 128, all powers of two, least padding; ties prefer wider rows
void pick_tile(int N, int Ho, int Wo, int* lw_out, int* lh_out) {
  int best_lw = 0, best_lh = 0; long long best = -1;
  for (int lw = 0; lw <= 7; ++lw)
    for (int lh = 0; lw + lh <= 7; ++lh) {
      const int bw = 1 << lw, bh = 1 << lh, bnn = 128 >> (lw + lh);
      long long tiles = (long long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * ((N + bnn - 1) / bnn);
      if (best < 0 || tiles < best || (tiles == best && lw > best_lw)) { best = tiles; best_lw = lw; best_lh = lh; }
    }
  *lw_out = best_lw; *lh_out = best_lh;
}

CUresult encode_weights(EncodeTiledFn enc, CUtensorMap* m, const void* w, long long K, int Cout, int box_k, int box_n,
                        CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_n};
  cuuint32_t es[2] = {1, 1};
  return enc(m, VTD_TMAP_16, 2, const_cast<void*>(w), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

CUresult encode_act4d(EncodeTiledFn enc, CUtensorMap* m, const void* base, int C, int W, int H, int N,
                      long long sW, long long sH, long long sN /*element strides*/, int bw, int bh, int bnn) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bnn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, VTD_TMAP_16, 4, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// output map of the TMA-store epilogue: box = the 32 consecutive tile rows of one epilogue warp x 128 bytes of channels
// up = 1: the map covers the half-resolution residual of an upsample-add; the box is the source pixels of the warp's box
// pool: the map covers the POOLED output (W, H already pooled); 1 = 2x2, 2 = (2,1)
CUresult encode_out(EncodeTiledFn enc, CUtensorMap* m, void* out, int out_f32, int C, int W, int H, int N, int lw, int lh, int up,
                    int* box_bytes, int pool = 0) {
  const int bw = 1 << lw, bh = 1 << lh;
  int sw = bw < 32 ? bw : 32;
  int sh = bh < 32 / sw ? bh : 32 / sw;
  const int sn = 32 / (sw * sh);
  if (up) { sw = sw > 1 ? sw / 2 : 1; sh = sh > 1 ? sh / 2 : 1; }
  if (pool == 1) { sw /= 2; sh /= 2; } else if (pool == 2) { sh /= 2; }
  if (box_bytes) *box_bytes = sw * sh * sn * 128;
  const cuuint64_t es = out_f32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)(128 / es), (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)sn};
  cuuint32_t est[4] = {1, 1, 1, 1};
  return enc(m, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : VTD_TMAP_16, 4, out, dims, strides, box, est,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

namespace tc {
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int sm_count() {                        // of the CURRENT device (cached per ordinal)
  static int sms[128] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 127;
  int v = __atomic_load_n(&sms[dev], __ATOMIC_RELAXED);
  if (!v) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    __atomic_store_n(&sms[dev], v, __ATOMIC_RELAXED);
  }
  return v;
}

}  // namespace tc

constexpr int SMEM_TOTAL = 227 * 1024;      // dynamic shared memory a CTA may use

struct TcPlan {
  TcMaps maps;
  TcParams p;
  HeadConsts hc;
  int block_n;
  int mode;
  int smem;
};

static void plan_finalize(TcPlan* pl);

bool tc_supported(const ConvDesc& d) {
  if (d.Cin % 64 != 0 || d.Cout % 64 != 0) return false;
  if (d.out_mode != OUT_NHWC) return false;
  if (d.stride != 1 && d.stride != 2) return false;
  if (d.stride == 2 && ((d.H & 1) || (d.W & 1))) return false;
  if (d.res_mode == RES_UP2 && ((d.Ho & 1) || (d.Wo & 1))) return false;
  return true;
}

// pool != 0: the tile is 16 px wide and at least 2 rows high, so that an epilogue warp's 32 pixels are a 16 x 2 box
// holding whole pooling windows
static void fill_common(TcPlan* pl, int N, int Ho, int Wo, int Cout, int Cin, int bn, int pool = 0) {
  TcParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.Cin = Cin;
  p.KH = p.KW = 1; p.stride = 1; p.pad = 0;
  pick_tile(N, Ho, Wo, &p.lw, &p.lh);
  if (pool) { p.lw = 4; p.lh = Ho >= 8 ? 3 : (Ho >= 4 ? 2 : 1); p.pool = pool; }
  if (const char* e = pool ? nullptr : dev_env("VTD_TILE")) {               // tuning aid: "lw,lh" for layers at least that large
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a + b <= 7 && (1 << a) <= Wo && (1 << b) <= Ho && (128 >> (a + b)) <= (N > 1 ? N : 1)) { p.lw = a; p.lh = b; }
  }
  const int bw = 1 << p.lw, bh = 1 << p.lh, bnn = 128 >> (p.lw + p.lh);
  p.tiles_x = (Wo + bw - 1) / bw; p.tiles_y = (Ho + bh - 1) / bh; p.tiles_n = (N + bnn - 1) / bnn;
  p.n_blocks = Cout / bn;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  pl->block_n = bn;
}

TcPlan* tc_plan_create(const ConvDesc& d, std::string* err) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  if (!tc_supported(d)) return fail("shape not supported by the tcgen05 path");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  pl->mode = MODE_CONV;
  const int bn = d.Cout % 256 == 0 ? 256 : (d.Cout % 128 == 0 ? 128 : 64);
  if (d.pool && (d.out_f32 || d.res_mode != RES_NONE || (d.Ho & 1) || (d.pool == 1 && (d.Wo & 1)) || d.N * d.Ho * d.Wo < 128)) {
    delete pl; return fail("fused pooling needs an even, bf16, residual-free output");
  }
  fill_common(pl, d.N, d.Ho, d.Wo, d.Cout, d.Cin, bn, d.pool);
  TcParams& p = pl->p;
  if (d.hint_lw >= 0 && d.hint_lh >= 0 && d.hint_lw + d.hint_lh <= 7 && !d.pool) {
    p.lw = d.hint_lw; p.lh = d.hint_lh;
    const int hbw = 1 << p.lw, hbh = 1 << p.lh, hbn = 128 >> (p.lw + p.lh);
    p.tiles_x = (d.Wo + hbw - 1) / hbw; p.tiles_y = (d.Ho + hbh - 1) / hbh; p.tiles_n = (d.N + hbn - 1) / hbn;
    p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  }
  p.KH = d.KH; p.KW = d.KW; p.stride = d.stride; p.pad = d.pad;
  p.relu = d.relu; p.res_mode = d.res_mode; p.out_f32 = d.out_f32;
  p.bias = d.bias; p.res = reinterpret_cast<const bf16*>(d.res); p.out = d.out;
  // 64 -> 64 3x3 s1 p1 (ResNet layer1): halo mode, see TcParams::halo.  Needs the weights resident (72 KB) and one 64-channel
  // chunk per pixel, tile 8 x 16 x 1 image.
  p.halo = (d.KH == 3 && d.KW == 3 && d.stride == 1 && d.pad == 1 && d.Cin == 64 && d.Cout == 64 && !d.pool && d.Ho >= 8 &&
            d.Wo >= 8 && !dev_env("VTD_NO_HALO") && !dev_env("VTD_NO_BRES")) ? 1 : 0;
  // any other 3x3 s1 p1 layer whose maps tile into 8 x 16 pixels with little waste: halo mode with streamed weights (2)
  // (a pooled layer only as a CTA pair, N = 128, whole pooling windows per 8 x 16 tile: the CRNN's 64 -> 128 layer)
  if (!p.halo && d.KH == 3 && d.KW == 3 && d.stride == 1 && d.pad == 1 && !dev_env("VTD_NO_HALO") &&
      (!d.pool || (bn == 128 && d.Cout == 128 && !d.out_f32 && d.N * ((d.Wo + 7) / 8) * ((d.Ho + 15) / 16) >= 2 &&
                   !dev_env("VTD_NO_POOL_HALO") && !dev_env("VTD_CTA2")))) {
    const long long covered = (long long)((d.Wo + 7) / 8 * 8) * ((d.Ho + 15) / 16 * 16);
    int mask = 3;                                         // bit 0: N = 128 layers, bit 1: N = 256 layers
    if (const char* e = dev_env("VTD_HALO2")) mask = atoi(e);
    const int waste = dev_env("VTD_HALO_WASTE") ? atoi(dev_env("VTD_HALO_WASTE")) : 13;   // percent of padded pixels tolerated
    if (covered * 100 <= (long long)d.Wo * d.Ho * (100 + waste) && ((bn == 128 && (mask & 1)) || (bn == 256 && (mask & 2)))) p.halo = 2;
  }
  if (p.halo) {
    p.lw = 3; p.lh = 4;
    p.tiles_x = (d.Wo + 7) / 8; p.tiles_y = (d.Ho + 15) / 16; p.tiles_n = d.N;
    p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
    CUresult hr = encode_act4d(enc, &pl->maps.a[1], d.in, d.Cin, d.W, d.H, d.N, d.Cin, (long long)d.W * d.Cin,
                               (long long)d.H * d.W * d.Cin, HALO_PW, HALO_PH, 1);
    if (hr != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(halo patch) failed: " + std::to_string((int)hr)); }
    if (p.halo == 2) {
      int cmask = 3;                                      // CTA pairs: bit 0 N = 128 layers, bit 1 N = 256 layers
      if (const char* e = dev_env("VTD_CTA2")) cmask = atoi(e);
      p.cta2 = (p.n_blocks == 1 && !d.out_f32 && p.total_tiles >= 2 && ((bn == 128 && (cmask & 1)) || (bn == 256 && (cmask & 2)))) ? 1 : 0;
      p.b_taps = (bn <= 128 || p.cta2) ? 3 : 1;
      const long long K = 9LL * d.Cin;
      cuuint64_t dims[4] = {64, (cuuint64_t)d.Cout, (cuuint64_t)(d.Cin / 64), 9};
      cuuint64_t strides[3] = {(cuuint64_t)K * 2, 128, (cuuint64_t)d.Cin * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)(p.cta2 ? bn / 2 : bn), 1, (cuuint32_t)p.b_taps};
      cuuint32_t es[4] = {1, 1, 1, 1};
      hr = enc(&pl->maps.b4, VTD_TMAP_16, 4, const_cast<void*>(d.w), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (hr != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weights by tap) failed: " + std::to_string((int)hr)); }
    }
  }
  const int bw = 1 << p.lw, bh = 1 << p.lh, bnn = 128 >> (p.lw + p.lh);
  const int nmaps = d.stride == 1 ? 1 : 4;
  for (int mi = 0; mi < nmaps; ++mi) {
    const int py = mi >> 1, px = mi & 1, st = d.stride;
    const char* base = reinterpret_cast<const char*>(d.in) + ((size_t)py * d.W + px) * d.Cin * 2;
    CUresult r = encode_act4d(enc, &pl->maps.a[mi], base, d.Cin, (d.W - px + st - 1) / st, (d.H - py + st - 1) / st, d.N,
                              (long long)d.Cin * st, (long long)d.W * d.Cin * st, (long long)d.H * d.W * d.Cin, bw, bh, bnn);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r)); }
  }
  CUresult r = encode_weights(enc, &pl->maps.b, d.w, (long long)d.KH * d.KW * d.Cin, d.Cout, 64, bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); }
  // CTA pairs for the remaining N = 256 layers with streamed weights (conv_tc2g_kernel): decided here, confirmed in plan_smem
  if (!p.halo && bn == 256 && !(d.out_f32 && d.res_mode != RES_NONE) &&
      (long long)d.KH * d.KW * (d.Cin / 64) * 256 * 128 > 96 * 1024 && p.tiles_x * p.tiles_y * p.tiles_n >= 2 &&
      !(dev_env("VTD_CTA2") && !(atoi(dev_env("VTD_CTA2")) & 4))) {
    p.cta2 = 2;
    r = encode_weights(enc, &pl->maps.b4, d.w, (long long)d.KH * d.KW * d.Cin, d.Cout, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weight halves) failed: " + std::to_string((int)r)); }
  }
  r = encode_out(enc, &pl->maps.o, d.out, d.out_f32, d.Cout, d.pool == 1 ? d.Wo / 2 : d.Wo, d.pool ? d.Ho / 2 : d.Ho, d.N, p.lw, p.lh, 0,
                 nullptr, d.pool);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(output) failed: " + std::to_string((int)r)); }
  if (d.res_mode != RES_NONE && !d.out_f32) {
    const int up = d.res_mode == RES_UP2 ? 1 : 0;
    r = encode_out(enc, &pl->maps.r, const_cast<void*>(d.res), 0, d.Cout, d.Wo >> up, d.Ho >> up, d.N, p.lw, p.lh, up, &p.res_bytes);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(residual) failed: " + std::to_string((int)r)); }
  }
  plan_finalize(pl);
  if (d.pool && !pl->p.epi_tma) { delete pl; return fail("no shared memory left for the fused-pool staging"); }
  return pl;
}

// Stem plan (MODE_WIN).  `in` points at a zero-bordered bf16 buffer [N][Hp][Wp][cpp] whose pixel (0,0) is the
// top-left corner of the receptive field of output (0,0); consecutive outputs are `stride` pixels apart and
// stride*cpp must be 8 elements (16 bytes).  Weights: [Cout=64][nr][32] bf16 (one 64-byte row per filter row).
TcPlan* tc_plan_create_win(const void* in, int N, int Hp, int Wp, int cpp, int stride, int nr, int Ho, int Wo,
                           const void* w, const float* bias, void* out, int relu, std::string* err, int pool) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  if (stride * cpp != 8 || (stride != 1 && stride != 2)) return fail("window stride must be 16 bytes");
  if ((Wo - 1) * stride * cpp + 32 > Wp * cpp) return fail("padded row too short for the last window");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  pl->mode = MODE_WIN;
  if (pool && ((Ho & 1) || (pool == 1 && (Wo & 1)))) { delete pl; return fail("fused pooling needs even output sizes"); }
  fill_common(pl, N, Ho, Wo, 64, 32, 64, pool);
  TcParams& p = pl->p;
  p.nr = nr; p.sdiv = stride; p.relu = relu; p.bias = bias; p.out = out;
  // DBNet stem: direct windows (see TcParams::win2).  One-row tiles of 128 outputs; the last tile of a row hangs over the
  // right edge (its copies run into the next padded row: the buffer has slack after the last image, see api.cu).
  p.win2 = (stride == 2 && cpp == 4 && nr == 7 && !pool && Wo >= 128 && !dev_env("VTD_NO_WIN2")) ? 1 : 0;
  if (p.win2) {
    p.lw = 7; p.lh = 0;
    p.tiles_x = (Wo + 127) / 128; p.tiles_y = Ho; p.tiles_n = N;
    p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
    p.win_in = reinterpret_cast<const uint8_t*>(in);
    p.win_rp = (long long)Wp * cpp * 2; p.win_ip = (long long)Hp * Wp * cpp * 2;
  }
  const int bw = 1 << p.lw, bh = 1 << p.lh, bnn = 128 >> (p.lw + p.lh);
  const long long row = (long long)Wp * cpp;            // elements per padded row
  // dims: window element, ox (16-byte stride: overlapping windows), row phase, oy, image
  cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)stride, (cuuint64_t)((Hp + stride - 1) / stride), (cuuint64_t)N};
  cuuint64_t strides[4] = {16, (cuuint64_t)row * 2, (cuuint64_t)row * stride * 2, (cuuint64_t)row * Hp * 2};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)bnn};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&pl->maps.a[0], VTD_TMAP_16, 5, const_cast<void*>(in), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(window) failed: " + std::to_string((int)r)); }
  r = encode_weights(enc, &pl->maps.b, w, (long long)nr * 32, 64, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); }
  r = encode_out(enc, &pl->maps.o, out, 0, 64, pool == 1 ? Wo / 2 : Wo, pool ? Ho / 2 : Ho, N, p.lw, p.lh, 0, nullptr, pool);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(output) failed: " + std::to_string((int)r)); }
  plan_finalize(pl);
  if (pool && !pl->p.epi_tma) { delete pl; return fail("no shared memory left for the fused-pool staging"); }
  return pl;
}

// DB head tail plan (MODE_DBHEAD).  feat: [N][H4][W4][128] bf16 (channels 0..63 probability branch, 64..127
// threshold branch); w1: [2][256][64] bf16.
TcPlan* tc_plan_create_dbhead(const void* feat, int N, int H4, int W4, const void* w1, const float* b1_host,
                              const float* w2_host, const float* b2_host, float* prob, float* thresh, uint8_t* mask,
                              std::string* err) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  pl->mode = MODE_DBHEAD;
  fill_common(pl, N, H4, W4, 512, 64, 256);
  TcParams& p = pl->p;
  p.prob = prob; p.thresh = thresh; p.mask = mask;
  memcpy(pl->hc.b1, b1_host, sizeof(pl->hc.b1));
  memcpy(pl->hc.w2, w2_host, sizeof(pl->hc.w2));
  memcpy(pl->hc.b2, b2_host, sizeof(pl->hc.b2));
  const int bw = 1 << p.lw, bh = 1 << p.lh, bnn = 128 >> (p.lw + p.lh);
  CUresult r = encode_act4d(enc, &pl->maps.a[0], feat, 128, W4, H4, N, 128, (long long)W4 * 128, (long long)H4 * W4 * 128,
                            bw, bh, bnn);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(head feature) failed: " + std::to_string((int)r)); }
  r = encode_weights(enc, &pl->maps.b, w1, 64, 512, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(head weights) failed: " + std::to_string((int)r)); }
  plan_finalize(pl);
  return pl;
}

// LSTM step plan (MODE_LSTM).  h_prev: [2 dirs][Bcap][256] bf16; whh: [2*1024][256] bf16, rows permuted to
// (dir, unit tile of 64, gate, unit).
TcPlan* tc_plan_create_lstm(const void* h_prev, void* h_next, int Bcap, const void* whh, const float* xproj, float* cbuf,
                            void* seq_out, int T, std::string* err) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  pl->mode = MODE_LSTM;
  fill_common(pl, 1, 1, Bcap, 2048, 256, 256);
  TcParams& p = pl->p;
  p.lw = 7; p.lh = 0;                                   // 128 sequences per tile
  p.tiles_x = (Bcap + 127) / 128; p.tiles_y = 1; p.tiles_n = 1;
  p.total_tiles = p.tiles_x * p.n_blocks;
  p.xproj = xproj; p.cbuf = cbuf; p.h_next = reinterpret_cast<bf16*>(h_next); p.seq_out = reinterpret_cast<bf16*>(seq_out);
  p.lstm_T = T; p.lstm_Bcap = Bcap;
  for (int dir = 0; dir < 2; ++dir) {
    const char* base = reinterpret_cast<const char*>(h_prev) + (size_t)dir * Bcap * 256 * 2;
    CUresult r = encode_act4d(enc, &pl->maps.a[dir], base, 256, Bcap, 1, 1, 256, (long long)Bcap * 256, (long long)Bcap * 256,
                              128, 1, 1);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(h) failed: " + std::to_string((int)r)); }
  }
  CUresult r = encode_weights(enc, &pl->maps.b, whh, 256, 2048, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(W_hh) failed: " + std::to_string((int)r)); }
  plan_finalize(pl);
  return pl;
}

// One-pass DB head plan (dbhead_fused_kernel).  d = the merged 3x3 256->128 convolution of the two branches (input P2,
// folded BN bias, ReLU); w1: [2][256][64] 16-bit (rows (dy,dx,co)), b1/w2/b2 host fp32 as for tc_plan_create_dbhead.
// Returns nullptr (err set) when the shape does not tile into 8 x 16 patches with little waste: the caller then keeps the
// two-kernel path (convolution -> feature map -> MODE_DBHEAD).
TcPlan* tc_plan_create_headfused(const ConvDesc& d, const void* w1, const float* b1_host, const float* w2_host,
                                 const float* b2_host, float* prob, float* thresh, uint8_t* mask, std::string* err) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  if (d.KH != 3 || d.KW != 3 || d.stride != 1 || d.pad != 1 || d.Cout != 128 || d.Cin % 64 != 0 || d.Ho < 8 || d.Wo < 8)
    return fail("not the merged 3x3 head convolution");
  const long long covered = (long long)((d.Wo + 7) / 8 * 8) * ((d.Ho + 15) / 16 * 16);
  if (covered * 100 > (long long)d.Wo * d.Ho * 113) return fail("map does not tile into 8 x 16 patches");
  if (dev_env("VTD_NO_HEAD_FUSION")) return fail("disabled");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  pl->mode = MODE_CONV;
  fill_common(pl, d.N, d.Ho, d.Wo, d.Cout, d.Cin, 128);
  TcParams& p = pl->p;
  p.KH = p.KW = 3; p.stride = 1; p.pad = 1; p.relu = 1; p.bias = d.bias;
  p.halo = 2; p.cta2 = 1; p.b_taps = 3; p.stages = 2; p.b_stages = 4;
  p.kps = 0;                                              // convolution groups in flight ahead of a tail MMA (0 = unlimited)
  if (const char* e = dev_env("VTD_HF_BST")) { int v = atoi(e); if (v >= 2 && v <= 4) p.b_stages = v; }
  if (const char* e = dev_env("VTD_HF_AHEAD")) { int v = atoi(e); if (v >= 0 && v <= p.b_stages) p.kps = v; }
  p.lw = 3; p.lh = 4;
  p.tiles_x = (d.Wo + 7) / 8; p.tiles_y = (d.Ho + 15) / 16; p.tiles_n = d.N; p.n_blocks = 1;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.prob = prob; p.thresh = thresh; p.mask = mask;
  if (p.total_tiles < 2) { delete pl; return fail("fewer than two tiles"); }
  memcpy(pl->hc.b1, b1_host, sizeof(pl->hc.b1));
  memcpy(pl->hc.w2, w2_host, sizeof(pl->hc.w2));
  memcpy(pl->hc.b2, b2_host, sizeof(pl->hc.b2));
  CUresult hr = encode_act4d(enc, &pl->maps.a[1], d.in, d.Cin, d.W, d.H, d.N, d.Cin, (long long)d.W * d.Cin,
                             (long long)d.H * d.W * d.Cin, HALO_PW, HALO_PH, 1);
  if (hr != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(halo patch) failed: " + std::to_string((int)hr)); }
  {
    const long long K = 9LL * d.Cin;
    cuuint64_t dims[4] = {64, (cuuint64_t)d.Cout, (cuuint64_t)(d.Cin / 64), 9};
    cuuint64_t strides[3] = {(cuuint64_t)K * 2, 128, (cuuint64_t)d.Cin * 2};
    cuuint32_t box[4] = {64, 64, 1, 3};
    cuuint32_t es[4] = {1, 1, 1, 1};
    hr = enc(&pl->maps.b4, VTD_TMAP_16, 4, const_cast<void*>(d.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (hr != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weights by tap) failed: " + std::to_string((int)hr)); }
  }
  hr = encode_weights(enc, &pl->maps.b, w1, 64, 512, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (hr != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(W1) failed: " + std::to_string((int)hr)); }
  pl->smem = p.stages * HALO_SLOT + p.b_stages * (3 * 64 * 128) + HF_A2_BYTES + HF_W1_BYTES + HF_CONST_BYTES + 512 + 1024;
  return pl;
}

cudaError_t dbhead_fused_tcgen05(TcPlan* pl, int n, float thr, const float* logit_bias, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  TcParams p = pl->p;
  p.N = n < pl->p.N ? n : pl->p.N;
  p.tiles_n = p.N;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.logit_bias = logit_bias;
  pl->hc.thr = thr;
  static PerDeviceFlag attr_done;
  cudaError_t e = once_per_device(attr_done, [] {
    return cudaFuncSetAttribute(dbhead_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (e != cudaSuccess) return e;
  const int pairs = (p.total_tiles + 1) / 2;
  const int clusters = pairs < sm_count() / 2 ? pairs : sm_count() / 2;
  dbhead_fused_kernel<<<2 * clusters, HF_THREADS, pl->smem, s>>>(pl->maps, p, pl->hc);
  if (lc) lc->n++;
  return cudaGetLastError();
}

// Fused stem + max-pool plan (stem_pool_kernel).  `in`: zero-bordered 16-bit input [N][dh + 6][dw + 10][4] (3 rows above /
// below, 6 px left, 4 px right of the image); w: window weights [64][7][8 taps][4 ch]; out: pooled map [N][dh/4][dw/4][64].
struct StemPoolPlan { CUtensorMap wmap; StemPoolParams p; int smem; bool crnn; };

StemPoolPlan* stem_pool_plan_create(const void* in, int N, int dh, int dw, const void* w, const float* bias, void* out, std::string* err) {
  auto fail = [&](const std::string& m) -> StemPoolPlan* { if (err) *err = m; return nullptr; };
  if (dh % 4 || dw % 4 || dw / 2 < 128) return fail("map too small for the fused stem");
  if (dev_env("VTD_NO_STEM_POOL")) return fail("disabled");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  StemPoolPlan* pl = new StemPoolPlan();
  memset(pl, 0, sizeof(*pl));
  StemPoolParams& p = pl->p;
  p.in = reinterpret_cast<const uint8_t*>(in);
  p.in_rp = (long long)(dw + 10) * 8; p.in_ip = p.in_rp * (dh + 6);
  p.bias = bias; p.out = reinterpret_cast<bf16*>(out);
  p.N = N; p.Ho = dh / 2; p.Wo = dw / 2; p.Hp = dh / 4; p.Wp = dw / 4;
  p.tiles_x = (p.Wp + 63 - 1) / 63;                       // 63 pooled columns per column tile (stem_pool_kernel<false>)
  // bands of pooled rows: enough items for ~4 per SM at the capacity batch, each band at least 8 rows
  int bands = (4 * sm_count() + N * p.tiles_x - 1) / (N * p.tiles_x);
  if (bands < 1) bands = 1;
  if (bands > p.Hp / 8) bands = p.Hp / 8 > 0 ? p.Hp / 8 : 1;
  p.rows_per_band = (p.Hp + bands - 1) / bands;
  p.bands = (p.Hp + p.rows_per_band - 1) / p.rows_per_band;
  p.total_items = N * p.tiles_x * p.bands;
  CUresult r = encode_weights(enc, &pl->wmap, w, 7 * 32, 64, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(stem weights) failed: " + std::to_string((int)r)); }
  const int fixed = 7 * 64 * 64 + 2 * SP_ROWBUF + 512 + 1024;
  int st = (SMEM_TOTAL - fixed) / WIN2_SLOT;
  p.stages = st > 8 ? 8 : st;
  pl->smem = p.stages * WIN2_SLOT + fixed;
  pl->crnn = false;
  return pl;
}

// The CRNN's first layer through the same kernel.  `in`: zero-bordered crops [N][34][cw + 4][8] (1 row above / below, 1 px
// left, 3 right); w: window weights [64][3][4 taps][8 ch]; out: pooled map [N][16][cw / 2][64].
StemPoolPlan* stem_pool_plan_create_crnn(const void* in, int N, int cw, const void* w, const float* bias, void* out, std::string* err) {
  auto fail = [&](const std::string& m) -> StemPoolPlan* { if (err) *err = m; return nullptr; };
  if (cw % 2 || cw < 16) return fail("crop width not supported by the fused first layer");
  if (dev_env("VTD_NO_STEM_POOL")) return fail("disabled");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  StemPoolPlan* pl = new StemPoolPlan();
  memset(pl, 0, sizeof(*pl));
  StemPoolParams& p = pl->p;
  p.in = reinterpret_cast<const uint8_t*>(in);
  p.in_rp = (long long)(cw + 4) * 16; p.in_ip = p.in_rp * 34;
  p.bias = bias; p.out = reinterpret_cast<bf16*>(out);
  p.N = N; p.Ho = 32; p.Wo = cw; p.Hp = 16; p.Wp = cw / 2;
  p.tiles_x = (cw + 127) / 128;
  p.bands = 1; p.rows_per_band = p.Hp;
  p.total_items = N * p.tiles_x;
  CUresult r = encode_weights(enc, &pl->wmap, w, 3 * 32, 64, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(CRNN stem weights) failed: " + std::to_string((int)r)); }
  const int fixed = 3 * 64 * 64 + 2 * SP_ROWBUF + 512 + 1024;
  int st = (SMEM_TOTAL - fixed) / (7 * 1024);
  p.stages = st > 12 ? 12 : st;
  pl->smem = p.stages * 7 * 1024 + fixed;
  pl->crnn = true;
  return pl;
}

void stem_pool_plan_destroy(StemPoolPlan* p) { delete p; }

cudaError_t stem_pool_tcgen05(const StemPoolPlan* pl, int n, cudaStream_t s, LaunchCounter* lc, const int* n_dyn, int n_first) {
  if (n <= 0) return cudaSuccess;
  StemPoolParams p = pl->p;
  p.N = n < pl->p.N ? n : pl->p.N;
  p.total_items = p.N * p.tiles_x * p.bands;
  p.n_dyn = n_dyn; p.n_first = n_first;
  static PerDeviceFlag attr_done[2];
  cudaError_t e = pl->crnn ? once_per_device(attr_done[1], [] {
    return cudaFuncSetAttribute(stem_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }) : once_per_device(attr_done[0], [] {
    return cudaFuncSetAttribute(stem_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (e != cudaSuccess) return e;
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  if (pl->crnn) stem_pool_kernel<true><<<grid, SP_THREADS, pl->smem, s>>>(pl->wmap, p);
  else stem_pool_kernel<false><<<grid, SP_THREADS, pl->smem, s>>>(pl->wmap, p);
  if (lc) lc->n++;
  return cudaGetLastError();
}

void tc_plan_destroy(TcPlan* p) { delete p; }


// ring depth / resident weights for a plan (called once per plan, after mode, block_n and p are filled)
template <int BN, int MODE>
static void plan_smem(TcPlan* pl) {
  using Cfg = TcCfg<BN, MODE>;
  TcParams& p = pl->p;
  const int ksteps = MODE == MODE_WIN ? p.nr : p.KH * p.KW * (p.Cin / BLOCK_K);
  const int bres_bytes = ksteps * Cfg::B_STAGE_BYTES;
  const bool can_res = (MODE == MODE_CONV || MODE == MODE_WIN) && p.n_blocks == 1 && bres_bytes <= 96 * 1024 &&
                       !dev_env("VTD_NO_BRES");
  p.bres = can_res ? 1 : 0;
  if (MODE == MODE_CONV && p.cta2 == 2) {                 // conv_tc2g_kernel: slots of A box + weight half, store staging
    p.bres = 0; p.kps = 1; p.res_tma = 0; p.epi_tma = 1; p.dbg = 0;
    const int slot = BLOCK_M * 128 + 128 * 128, fixed2 = 1024 + 256;
    const int stg2 = NUM_EPI_WARPS * 4096 + (p.pool ? NUM_EPI_WARPS * 2048 : 0);      // + pooled-box staging
    int st = (SMEM_TOTAL - fixed2 - stg2) / slot;
    p.stages = st > 8 ? 8 : st;
    pl->smem = p.stages * slot + stg2 + fixed2;
    return;
  }
  if (MODE == MODE_CONV && p.halo == 2) {
    // patch ring (2 slots: a patch lasts 36 MMAs, one ahead is enough) + weight ring (what is left, 2..8 slots of b_taps
    // taps) + store staging
    p.bres = 0; p.kps = 1; p.res_tma = 0;
    if (p.cta2) {                                         // conv_tc2_kernel: patch ring (2) + this CTA's weight halves + staging
      const int bslot2 = p.b_taps * (BN / 2) * 128;
      const int fixed2 = 1024 + 512;
      p.stages = 2; p.epi_tma = 1;
      const int stg2 = NUM_EPI_WARPS * 4096 + (p.pool ? NUM_EPI_WARPS * 2048 : 0);
      int bst = (SMEM_TOTAL - fixed2 - p.stages * HALO_SLOT - stg2) / bslot2;
      p.b_stages = bst > 8 ? 8 : bst;
      // one 64-channel chunk and room for all nine taps: the weights stay (conv_tc2_kernel, p.bres), the patch ring deepens
      if (p.Cin == 64 && p.b_taps == 3 && bst >= 3 && !dev_env("VTD_NO_BRES")) {
        p.bres = 1; p.b_stages = 3;
        int ast = (SMEM_TOTAL - fixed2 - 3 * bslot2 - stg2) / HALO_SLOT;
        p.stages = ast > 4 ? 4 : ast;
      }
      p.dbg = 0;
      pl->smem = p.stages * HALO_SLOT + p.b_stages * bslot2 + stg2 + fixed2;
      return;
    }
    const int stg = NUM_EPI_WARPS * 4096;
    const int bslot = p.b_taps * Cfg::B_STAGE_BYTES;
    const int fixed = 1024 + Cfg::TAIL_BYTES;
    p.stages = 2;
    p.epi_tma = (p.out_f32 && p.res_mode != RES_NONE) || dev_env("VTD_NO_TMA_STORE") ? 0 : 1;
    int bst = (SMEM_TOTAL - fixed - p.stages * HALO_SLOT - (p.epi_tma ? stg : 0)) / bslot;
    if (bst < 3 && p.epi_tma && BN == 256) { p.epi_tma = 0; bst = (SMEM_TOTAL - fixed - p.stages * HALO_SLOT) / bslot; }
    if (bst > 8) bst = 8;
    if (bst < 2) { p.halo = 0; }                          // (cannot happen for N <= 256)
    else {
      p.b_stages = bst;
      // the residual by TMA as well when its staging still fits
      if (p.epi_tma && p.res_mode != RES_NONE && !dev_env("VTD_NO_TMA_RES") &&
          (SMEM_TOTAL - fixed - p.stages * HALO_SLOT - 2 * stg) / bslot >= 2) {
        p.res_tma = 1;
        p.b_stages = (SMEM_TOTAL - fixed - p.stages * HALO_SLOT - 2 * stg) / bslot;
        if (p.b_stages > 8) p.b_stages = 8;
      }
      p.dbg = 0;
      pl->smem = p.stages * HALO_SLOT + p.b_stages * bslot + (p.epi_tma ? stg : 0) + (p.res_tma ? stg : 0) + fixed;
      return;
    }
  }
  const int step_bytes = p.bres ? Cfg::A_BYTES : Cfg::STAGE_BYTES;
  // K steps per ring slot.  A slot handshake costs ~300-500 cycles in the two role warps; the MMAs of one K step take
  // 4 x max(32, N/2) cycles.  N = 256 hides it with one step per slot; narrower tiles batch up to 3 steps (a divisor of
  // the step count, at least 3 slots in the ring); the 3-channel stems (2 MMAs per filter row) take the whole filter.
  int want = MODE == MODE_WIN ? 7 : (BN >= 256 || p.halo ? 1 : 3);
  if (MODE == MODE_LSTM || MODE == MODE_DBHEAD) want = 1;
  if (const char* e = dev_env("VTD_KPS")) { int v = atoi(e); if (v >= 1 && BN < 256) want = v; }
  auto pick = [&](int avail, int* stages_out) {
    int kps = 1;
    for (int c = want; c >= 1; --c) {
      if (MODE == MODE_WIN ? (c != 7 && c != 3 && c != 1) : c > 3) continue;       // instantiated variants
      if (ksteps % c == 0 && avail / (p.win2 ? WIN2_SLOT : p.halo ? HALO_SLOT : c * step_bytes) >= (c == 1 ? 2 : 3)) { kps = c; break; }
    }
    int st = avail / (p.win2 ? WIN2_SLOT : p.halo ? HALO_SLOT : kps * step_bytes);
    if (st > Cfg::MAX_STAGES) st = Cfg::MAX_STAGES;
    if (!p.bres && kps == 1 && st > Cfg::STAGES) st = Cfg::STAGES;
    *stages_out = st;
    return kps;
  };
  const int avail = SMEM_TOTAL - (p.bres ? bres_bytes : 0) - 1024 - Cfg::TAIL_BYTES;
  const int stg_bytes = NUM_EPI_WARPS * 4096;
  const int pool_bytes = p.pool ? NUM_EPI_WARPS * 2048 : 0;
  int st_plain = 0, st_tma = 0;
  const int kps_plain = pick(avail, &st_plain);
  const int kps_tma = pick(avail - stg_bytes - pool_bytes, &st_tma);
  // the TMA-store epilogue needs 32 KB of staging: take it unless that costs K-step batching or leaves < 3 ring slots
  // (the resident-weight 64-channel 3x3 layers, where the handshake per K step is the larger cost)
  bool tma = (MODE == MODE_CONV || MODE == MODE_WIN) && !dev_env("VTD_NO_TMA_STORE") && kps_tma == kps_plain &&
             st_tma >= (st_plain < 3 ? st_plain : 3);
  if (p.pool && !tma) { p.epi_tma = 0; p.kps = kps_plain; p.stages = st_plain; p.res_tma = 0; p.dbg = 0; pl->smem = 0; return; }   // caller falls back
  if (BN == 256 && tma && st_tma < 4 && st_plain >= 4) tma = false;
  if (p.out_f32 && p.res_mode != RES_NONE) tma = false;     // (no such layer) residual registers are sized for bf16 groups
  p.epi_tma = tma ? 1 : 0;
  p.kps = tma ? kps_tma : kps_plain;
  p.stages = tma ? st_tma : st_plain;
  // residual through TMA as well when another 32 KB leave the ring as deep (and the warp's box is at least 2 px wide
  // for the upsample-add)
  p.res_tma = 0;
  if (tma && p.res_mode != RES_NONE && !dev_env("VTD_NO_TMA_RES") && (p.res_mode != RES_UP2 || p.lw >= 1)) {
    int st_r = 0;
    if (pick(avail - 2 * stg_bytes - pool_bytes, &st_r) == p.kps && st_r >= (p.halo ? 3 : (p.stages < 4 ? p.stages : 4))) { p.res_tma = 1; p.stages = st_r; }
  }
  if (const char* e = dev_env("VTD_TC_STAGES")) {          // tuning aid: cap the ring depth
    int cap = atoi(e);
    if (cap >= 2 && cap < p.stages) p.stages = cap;
  }
  p.dbg = dev_env("VTD_DBG") ? atoi(dev_env("VTD_DBG")) : 0;
  pl->smem = p.stages * (p.win2 ? WIN2_SLOT : p.halo ? HALO_SLOT : p.kps * step_bytes) + (p.bres ? bres_bytes : 0) + (tma ? stg_bytes + pool_bytes : 0) + (p.res_tma ? stg_bytes : 0) + 1024 +
             Cfg::TAIL_BYTES;
}

static void plan_finalize(TcPlan* pl) {
  switch (pl->mode) {
    case MODE_WIN: plan_smem<64, MODE_WIN>(pl); break;
    case MODE_DBHEAD: plan_smem<256, MODE_DBHEAD>(pl); break;
    case MODE_LSTM: plan_smem<256, MODE_LSTM>(pl); break;
    default:
      if (pl->block_n == 256) plan_smem<256, MODE_CONV>(pl);
      else if (pl->block_n == 128) plan_smem<128, MODE_CONV>(pl);
      else plan_smem<64, MODE_CONV>(pl);
  }
}

template <int BN, int MODE, int KPS = 1>
static cudaError_t launch_tc(const TcPlan* pl, const TcParams& p, const typename ExtraOf<MODE>::type& ex, cudaStream_t s,
                             bool pdl = false) {
  static PerDeviceFlag attr_done;
  {
    cudaError_t e = once_per_device(attr_done, [] {
      return cudaFuncSetAttribute(conv_tc_kernel<BN, MODE, KPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (e != cudaSuccess) return e;
  }
  const int sms = sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = (size_t)pl->smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
#ifdef VTD_TIMERS
  if (getenv("VTD_TIMERS")) {
    static long long* tb = nullptr;
    if (!tb) cudaMalloc(&tb, 148 * 9 * sizeof(long long));
    TcParams q = p; q.timers = tb;
    cudaMemsetAsync(tb, 0, 148 * 9 * sizeof(long long), s);
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, MODE, KPS>, pl->maps, q, ex);
    cudaStreamSynchronize(s);
    static long long h[148 * 9];
    cudaMemcpy(h, tb, sizeof(h), cudaMemcpyDeviceToHost);
    double a[9] = {0};
    for (int i = 0; i < grid; ++i) for (int j = 0; j < 9; ++j) a[j] += (double)h[i * 9 + j] / grid;
    const int tiles_cta = (p.total_tiles + grid - 1) / grid;
    fprintf(stderr, "TMR BN=%d mode=%d kps=%d st=%d %dx%d %d->%d k%d s%d tiles/cta=%d | prod tot %.0f wait_empty %.0f | mma tot %.0f "
            "wait_full %.0f wait_tempty %.0f | epi tot %.0f wait_tfull %.0f | per tile %.0f\n", BN, MODE, KPS, p.stages, p.Ho, p.Wo,
            p.Cin, p.Cout, p.KH, p.stride, tiles_cta, a[0], a[1], a[3], a[4], a[5], a[6], a[7], a[3] / tiles_cta);
    return e;
  }
#endif
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, MODE, KPS>, pl->maps, p, ex);
}

// n_actual: images (MODE_LSTM: sequences) actually present in this call (<= what the plan was built for)
cudaError_t conv_tcgen05(const TcPlan* pl, int n_actual, cudaStream_t s, LaunchCounter* lc, const int* n_dyn, int n_first) {
  if (n_actual <= 0) return cudaSuccess;
  TcParams p = pl->p;
  p.n_dyn = n_dyn; p.n_first = n_first;
  const int bnn = 128 >> (p.lw + p.lh);
  p.N = n_actual < pl->p.N ? n_actual : pl->p.N;
  p.tiles_n = (p.N + bnn - 1) / bnn;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  NoExtra none{0};
  cudaError_t e;
  if (pl->mode == MODE_CONV && p.cta2 == 2) {
    static PerDeviceFlag attrg_done;
    {
      cudaError_t ae = once_per_device(attrg_done, [] {
        return cudaFuncSetAttribute(conv_tc2g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      });
      if (ae != cudaSuccess) return ae;
    }
    const int spatial = p.tiles_x * p.tiles_y * p.tiles_n;
    const int pairs = ((spatial + 1) / 2) * p.n_blocks;
    const int clusters = pairs < sm_count() / 2 ? pairs : sm_count() / 2;
    conv_tc2g_kernel<<<2 * clusters, NUM_THREADS, pl->smem, s>>>(pl->maps, p);
    if (lc) lc->n++;
    return cudaGetLastError();
  }
  if (pl->mode == MODE_CONV && p.cta2) {
    static PerDeviceFlag attr2_done[2];
    const int which = pl->block_n == 256 ? 1 : 0;
    {
      cudaError_t ae = once_per_device(attr2_done[which], [which] {
        return which ? cudaFuncSetAttribute(conv_tc2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                     : cudaFuncSetAttribute(conv_tc2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      });
      if (ae != cudaSuccess) return ae;
    }
    const int pairs = (p.total_tiles + 1) / 2;
    const int clusters = pairs < sm_count() / 2 ? pairs : sm_count() / 2;
    if (which) conv_tc2_kernel<256><<<2 * clusters, NUM_THREADS, pl->smem, s>>>(pl->maps, p);
    else conv_tc2_kernel<128><<<2 * clusters, NUM_THREADS, pl->smem, s>>>(pl->maps, p);
    if (lc) lc->n++;
    return cudaGetLastError();
  }
  const int key = pl->mode == MODE_WIN ? 1000 + p.kps : pl->block_n * 10 + p.kps;
  switch (key) {                                  // the (tile width, K steps per slot) pairs plan_smem() can pick
    case 1007: e = launch_tc<64, MODE_WIN, 7>(pl, p, none, s); break;
    case 1003: e = launch_tc<64, MODE_WIN, 3>(pl, p, none, s); break;
    case 1001: e = launch_tc<64, MODE_WIN, 1>(pl, p, none, s); break;
    case 2561: e = launch_tc<256, MODE_CONV, 1>(pl, p, none, s); break;
    case 1281: e = launch_tc<128, MODE_CONV, 1>(pl, p, none, s); break;
    case 1282: e = launch_tc<128, MODE_CONV, 2>(pl, p, none, s); break;
    case 1283: e = launch_tc<128, MODE_CONV, 3>(pl, p, none, s); break;
    case 641: e = launch_tc<64, MODE_CONV, 1>(pl, p, none, s); break;
    case 642: e = launch_tc<64, MODE_CONV, 2>(pl, p, none, s); break;
    case 643: e = launch_tc<64, MODE_CONV, 3>(pl, p, none, s); break;
    default: return cudaErrorInvalidConfiguration;
  }
  if (lc) lc->n++;
  return e;
}

cudaError_t dbhead_tcgen05(TcPlan* pl, int n, float thr, const float* logit_bias, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return cudaSuccess;
  TcParams p = pl->p;
  const int bnn = 128 >> (p.lw + p.lh);
  p.N = n < pl->p.N ? n : pl->p.N;
  p.tiles_n = (p.N + bnn - 1) / bnn;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  p.logit_bias = logit_bias;
  pl->hc.thr = thr;
  cudaError_t e = launch_tc<256, MODE_DBHEAD>(pl, p, pl->hc, s);
  if (lc) lc->n++;
  return e;
}

cudaError_t lstm_step_tcgen05(const TcPlan* pl, int B, int step, cudaStream_t s, LaunchCounter* lc) {
  if (B <= 0) return cudaSuccess;
  TcParams p = pl->p;
  p.lstm_B = B; p.lstm_step = step;
  p.Wo = B;                                              // rows beyond B are masked in the epilogue
  p.tiles_x = (B + 127) / 128;
  p.total_tiles = p.tiles_x * p.n_blocks;
  NoExtra none{0};
  cudaError_t e = launch_tc<256, MODE_LSTM>(pl, p, none, s, /*pdl=*/step > 0);
  if (lc) lc->n++;
  return e;
}

}  // namespace vtd
