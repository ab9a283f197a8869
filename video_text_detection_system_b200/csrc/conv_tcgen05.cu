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
// CTA = 6 warps, persistent over output tiles (static stride schedule):
//   warp 0 lane 0 : TMA producer            (full/empty mbarrier ring of STAGES slots)
//   warp 1        : TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma 128 x BLOCK_N x 16,
//                   accumulators in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i
//                   overlaps the MMAs of tile i+1; tcgen05.commit releases smem slots / publishes the tile
//   warps 2..5    : epilogue: tcgen05.ld 32 lanes x 32 columns -> +bias (folded BN) -> +residual (same size,
//                   or nearest-2x upsampled = the FPN top-down add) -> ReLU -> bf16 (or fp32) NHWC store.
// Every mbarrier wait is bounded (clock64) and traps instead of hanging the GPU.
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace vtd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB

struct alignas(64) TcMaps {
  CUtensorMap a[4];     // activation maps; [0] only for stride 1, [py*2+px] for stride 2
  CUtensorMap b;        // weights
};

struct TcParams {
  int N, Ho, Wo, Cout, Cin;
  int KH, KW, stride, pad;
  int lw, lh;                 // log2(BW), log2(BH); BN = 128 >> (lw+lh)
  int tiles_x, tiles_y, tiles_n, n_blocks;
  long long total_tiles;
  int relu, res_mode, out_f32;
  const float* bias;
  const bf16* res;
  void* out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins == 64) t0 = clock64();
    if (spins > 64 && (clock64() - t0) > 4000000000LL) __trap();   // ~2 s: fail loudly, never hang the box
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);      // start address, 16-byte units
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N=BLOCK_N
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

template <int BLOCK_N>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;     // 128 / 256 / 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms), barriers after it
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + Cfg::STAGES * Cfg::STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars);                       // [STAGES]
  const uint32_t empty0 = full0 + 8 * Cfg::STAGES;             // [STAGES]
  const uint32_t tfull0 = empty0 + 8 * Cfg::STAGES;            // [2]
  const uint32_t tempty0 = tfull0 + 16;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
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

  const int BW = 1 << p.lw, BH = 1 << p.lh;
  const int BNt = BLOCK_M >> (p.lw + p.lh);
  const int kchunks = p.Cin / BLOCK_K;
  const int ksteps = p.KH * p.KW * kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        long long t = tile;
        const int nb = (int)(t % p.n_blocks); t /= p.n_blocks;
        const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
        const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
        const int x0 = tx * BW, y0 = ty * BH, n0 = (int)t * BNt;
        for (int r = 0; r < p.KH; ++r) {
          for (int s = 0; s < p.KW; ++s) {
            int mi = 0, xo, yo;
            if (p.stride == 1) { xo = s - p.pad; yo = r - p.pad; }
            else {
              const int tyy = r - p.pad, txx = s - p.pad;
              const int py = tyy & 1, px = txx & 1;
              mi = py * 2 + px; yo = (tyy - py) / 2; xo = (txx - px) / 2;
            }
            const int kbase = (r * p.KW + s) * p.Cin;
            for (int kc = 0; kc < kchunks; ++kc) {
              mbar_wait(empty0 + 8 * stage, phase ^ 1);
              const uint32_t sa = ring + stage * Cfg::STAGE_BYTES;
              const uint32_t sb = sa + A_STAGE_BYTES;
              const uint32_t fb = full0 + 8 * stage;
              mbar_expect_tx(fb, Cfg::STAGE_BYTES);
              tma_load_4d(sa, &maps.a[mi], fb, kc * BLOCK_K, x0 + xo, y0 + yo, n0);
              tma_load_2d(sb, &maps.b, fb, kbase + kc * BLOCK_K, nb * BLOCK_N);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = (int)(it & 1);
        mbar_wait(tempty0 + 8 * as, (uint32_t)((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full0 + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = ring + stage * Cfg::STAGE_BYTES;
          const uint64_t ad = umma_desc(sa), bd = umma_desc(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ks | k) ? 1u : 0u);
          umma_commit(empty0 + 8 * stage);              // frees the slot when these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * as);                   // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;                        // row of the tile = pixel
    const int xx = m & (BW - 1), yy = (m >> p.lw) & (BH - 1), nn = m >> (p.lw + p.lh);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      long long t = tile;
      const int nb = (int)(t % p.n_blocks); t /= p.n_blocks;
      const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
      const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
      const int ox = tx * BW + xx, oy = ty * BH + yy, n = (int)t * BNt + nn;
      const bool valid = ox < p.Wo && oy < p.Ho && n < p.N;
      const int as = (int)(it & 1);
      mbar_wait(tfull0 + 8 * as, (uint32_t)((it >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const size_t opix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
      size_t rpix = 0;
      if (p.res_mode == RES_SAME) rpix = opix;
      else if (p.res_mode == RES_UP2) rpix = ((size_t)n * (p.Ho >> 1) + (oy >> 1)) * (p.Wo >> 1) + (ox >> 1);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c0), v);
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
          if (p.res_mode != RES_NONE) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + rpix * p.Cout + co);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u = __ldg(rp + j);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 r2 = __bfloat1622float2(h[e]);
                f[j * 8 + e * 2] += r2.x; f[j * 8 + e * 2 + 1] += r2.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
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
              __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
              u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
              u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
              op[j] = u;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    }
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

// ---- host side ---------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

}  // namespace

struct TcPlan {
  TcMaps maps;
  TcParams p;
  int block_n;
  int grid;
};

bool tc_supported(const ConvDesc& d) {
  if (d.Cin % 64 != 0 || d.Cout % 64 != 0) return false;
  if (d.out_mode != OUT_NHWC) return false;
  if (d.stride != 1 && d.stride != 2) return false;
  if (d.stride == 2 && ((d.H & 1) || (d.W & 1))) return false;
  if (d.res_mode == RES_UP2 && ((d.Ho & 1) || (d.Wo & 1))) return false;
  return true;
}

TcPlan* tc_plan_create(const ConvDesc& d, std::string* err) {
  auto fail = [&](const std::string& m) -> TcPlan* { if (err) *err = m; return nullptr; };
  if (!tc_supported(d)) return fail("shape not supported by the tcgen05 path");
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled not available from the driver");
  TcPlan* pl = new TcPlan();
  memset(&pl->maps, 0, sizeof(pl->maps));
  const int bn = d.Cout % 256 == 0 ? 256 : (d.Cout % 128 == 0 ? 128 : 64);
  pl->block_n = bn;
  // tile shape: BN x BH x BW = 128, all powers of two, least padding
  int best_lw = 0, best_lh = 0; long long best = -1;
  for (int lw = 0; lw <= 7; ++lw)
    for (int lh = 0; lw + lh <= 7; ++lh) {
      const int bw = 1 << lw, bh = 1 << lh, bnn = 128 >> (lw + lh);
      long long tiles = (long long)((d.Wo + bw - 1) / bw) * ((d.Ho + bh - 1) / bh) * ((d.N + bnn - 1) / bnn);
      // tie-break: prefer wider rows (longer contiguous stores / TMA lines)
      if (best < 0 || tiles < best || (tiles == best && lw > best_lw)) { best = tiles; best_lw = lw; best_lh = lh; }
    }
  TcParams& p = pl->p;
  p.N = d.N; p.Ho = d.Ho; p.Wo = d.Wo; p.Cout = d.Cout; p.Cin = d.Cin;
  p.KH = d.KH; p.KW = d.KW; p.stride = d.stride; p.pad = d.pad;
  p.lw = best_lw; p.lh = best_lh;
  const int bw = 1 << p.lw, bh = 1 << p.lh, bnn = 128 >> (p.lw + p.lh);
  p.tiles_x = (d.Wo + bw - 1) / bw; p.tiles_y = (d.Ho + bh - 1) / bh; p.tiles_n = (d.N + bnn - 1) / bnn;
  p.n_blocks = d.Cout / bn;
  p.total_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  p.relu = d.relu; p.res_mode = d.res_mode; p.out_f32 = d.out_f32;
  p.bias = d.bias; p.res = reinterpret_cast<const bf16*>(d.res); p.out = d.out;
  // activation map(s)
  const int nmaps = d.stride == 1 ? 1 : 4;
  for (int mi = 0; mi < nmaps; ++mi) {
    const int py = mi >> 1, px = mi & 1;
    const int st = d.stride;
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)((d.W - px + st - 1) / st),
                          (cuuint64_t)((d.H - py + st - 1) / st), (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)d.Cin * 2 * st, (cuuint64_t)d.W * d.Cin * 2 * st,
                             (cuuint64_t)d.H * d.W * d.Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bnn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    char* base = reinterpret_cast<char*>(const_cast<void*>(d.in)) + ((size_t)py * d.W + px) * d.Cin * 2;
    CUresult r = enc(&pl->maps.a[mi], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r)); }
  }
  {
    const long long K = (long long)d.KH * d.KW * d.Cin;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)d.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&pl->maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; return fail("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pl->grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  return pl;
}

void tc_plan_destroy(TcPlan* p) { delete p; }

template <int BN>
static cudaError_t launch_tc(const TcPlan* pl, const TcParams& p, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TcCfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  conv_tc_kernel<BN><<<pl->grid, NUM_THREADS, TcCfg<BN>::SMEM_BYTES, s>>>(pl->maps, p);
  return cudaGetLastError();
}

// n_actual: images actually present in this call (<= the N the plan was built for)
cudaError_t conv_tcgen05(const TcPlan* pl, int n_actual, cudaStream_t s, LaunchCounter* lc) {
  if (n_actual <= 0) return cudaSuccess;
  TcParams p = pl->p;
  const int bnn = 128 >> (p.lw + p.lh);
  p.N = n_actual < pl->p.N ? n_actual : pl->p.N;
  p.tiles_n = (p.N + bnn - 1) / bnn;
  p.total_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.n_blocks;
  TcPlan tmp = *pl;   // grid for the reduced tile count
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  tmp.grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  cudaError_t e;
  switch (pl->block_n) {
    case 256: e = launch_tc<256>(&tmp, p, s); break;
    case 128: e = launch_tc<128>(&tmp, p, s); break;
    default: e = launch_tc<64>(&tmp, p, s); break;
  }
  if (lc) lc->n++;
  return e;
}

}  // namespace vtd
