// Persistent BiLSTM layer on tcgen05 (bf16 tier): ONE launch runs all T timesteps of both directions.
//
// Replaces nn.LSTM(512, 256, num_layers=2, bidirectional=True, batch_first=True) of the CRNN recogniser
// (text_recognizer.py:26), one layer per launch; the input projection W_ih x + b_ih + b_hh of all timesteps is a
// separate GEMM (conv_tcgen05.cu) that leaves xproj [B][T][2][1024] fp32 with the gate rows permuted to
// (unit tile of 64, gate, unit).
//
// Work split: a thread-block CLUSTER of 4 CTAs owns 128 sequences of one direction; CTA j of the cluster owns hidden
// units 64j..64j+63 (all four gates = 256 gate rows).  Per CTA, for the whole kernel:
//   * its W_hh slice [256 gate rows x 256] bf16 (128 KB) is loaded once by TMA and stays in shared memory;
//   * h_{t-1} of the 128 sequences (all 256 units, bf16, K-major SWIZZLE_128B = the tcgen05 A operand, 64 KB) lives in
//     shared memory.  After a step every CTA writes its 64-unit slice of h_t straight into the A buffers of all four
//     CTAs of the cluster through distributed shared memory (st.shared::cluster) -- h never goes to global memory
//     between steps, and there is no TMA on the recurrent path;
//   * the cell state c stays in registers (each epilogue thread owns one sequence x 32 units for all timesteps).
// Per step: 16 tcgen05.mma (128 x 256 x 16) into TMEM -> epilogue warps add xproj, apply the gates, write h (DSMEM)
// and the layer output (global).  Ordering is carried by mbarriers only: a multicast tcgen05.commit tells all four
// CTAs "my MMAs no longer read h_{t-1}" (afree, 4 arrivals); h_t travels with st.async, whose bytes complete the
// transaction count of the destination CTA's hready barrier, which that CTA's MMA warp waits on (expect_tx 64 KB).
// No fences, no cluster-wide barriers inside the loop.  (Earlier versions: barrier.cluster pairs -- 12.7 us/step;
// st.shared::cluster + release/acquire.cluster mbarrier ops -- 10 us/step, MEMBAR.ALL.GPU on every arrive and
// CCTL.IVALL on every poll.)
#include "common.cuh"
#include "tc_common.cuh"
#include <string>
#include <stdio.h>
#include <stdlib.h>

namespace vtd {
namespace {

using namespace tc;

#ifndef VTD_LSTM_PARTS
#define VTD_LSTM_PARTS 4
#endif
constexpr int L_PARTS = VTD_LSTM_PARTS;        // epilogue warps per TMEM lane quarter: each owns 64 / L_PARTS hidden units of a row
constexpr int L_UPT = 64 / L_PARTS;            // hidden units per epilogue thread
constexpr int L_CH = L_UPT / 8;                // 16-byte chunks (8 bf16) per thread and gate
constexpr int L_EPI_THREADS = 128 * L_PARTS;
constexpr int L_THREADS = 64 + L_EPI_THREADS;  // warp 0: weight loader, warp 1: MMA issuer, then the epilogue warps
constexpr int L_W_BYTES = 4 * 256 * 128;       // 4 K-chunks x 256 rows x 128 B
constexpr int L_A_BYTES = 4 * 128 * 128;       // 4 K-chunks x 128 rows x 128 B
constexpr int L_SMEM = L_W_BYTES + L_A_BYTES + 1024 + 64;

#ifdef VTD_TIMERS
#define LT_DECL long long lt_a = 0, lt_b = 0, lt_c = 0, lt_d = 0, lt_e = 0, lt_f = 0; const long long lt_t0 = clock64();
#define LT(acc, stmt) { const long long _a = clock64(); stmt; acc += clock64() - _a; }
#else
#define LT_DECL
#define LT(acc, stmt) { stmt; }
#endif
struct LstmParams {
  long long* timers;
  const bf16* xproj;      // [B][T][2][1024] bf16
  bf16* seq_out;          // [B][T][512]
  int B, T;
  int rows;               // sequences per cluster: 128, or 64 when that still fits one wave (halves the h exchange per step)
  const int* n_dyn; int n_first;   // optional: the number of sequences lives on the device (clamp(*n_dyn - n_first, 0, B))
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {      // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope, bounded
  uint32_t done = 0;
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins == 64) t0 = clock64();
    if (spins > 64 && (clock64() - t0) > 4000000000LL) __trap();
  }
}
// tcgen05.commit that arrives on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// 16-byte store into (possibly remote) shared memory of the cluster through the ASYNC proxy, completing 16 bytes of
// the transaction count of the mbarrier `mbar` in the same destination CTA.  No fences on either side: the consumer's
// mbarrier wait orders the data before its tcgen05.mma reads.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(L_THREADS, 1)
bilstm_persistent_kernel(const __grid_constant__ CUtensorMap map_whh, const LstmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t wbuf = base;                           // [4][256 rows][128 B]
  const uint32_t abuf = base + L_W_BYTES;               // [4][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + L_W_BYTES + L_A_BYTES);
  const uint32_t wfull = smem_u32(bars), tfull = wfull + 8;
  const uint32_t afree = wfull + 16;      // 4 arrivals: the MMAs of a step have completed in every CTA of the cluster
  const uint32_t hready = wfull + 24;     // [4], one per K-chunk = per source CTA: the chunk of h_t written by CTA kc has landed
                                          // (own chunk: the local epilogue's arrive; a peer's: expect_tx of this MMA thread + its bulk copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x;                            // == rank in cluster: hidden units 64*jt ..
  const int dir = blockIdx.y;
  const int b0 = blockIdx.z * p.rows;
  int B = p.B;
  if (p.n_dyn) {
    const int c = *p.n_dyn - p.n_first;
    B = c < 0 ? 0 : (c > p.B ? p.B : c);
  }
  if (b0 >= B) return;                                  // the whole cluster (same blockIdx.z) leaves before any barrier exists

  if (warp == 0 && lane == 0) {
    mbar_init(wfull, 1);
    mbar_init(tfull, 1);
    mbar_init(afree, 4);
    for (int kc = 0; kc < 4; ++kc) mbar_init(hready + 8 * kc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whh) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // h_0 = 0: clear the A operand
  for (int i = threadIdx.x; i < L_A_BYTES / 16; i += L_THREADS)
    *reinterpret_cast<uint4*>(base_ptr + L_W_BYTES + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy writes -> visible to the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;
  cluster_arrive();
  cluster_wait();                                       // every CTA of the cluster is initialised

  if (warp == 0) {
    // ---- weights: once
    if (elect_one()) {
      mbar_expect_tx(wfull, L_W_BYTES);
      for (int kc = 0; kc < 4; ++kc)
        tma_load_2d(wbuf + kc * (256 * 128), &map_whh, wfull, kc * 64, (dir * 4 + jt) * 256);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = umma_idesc(256);
    mbar_wait(wfull, 0);
    LT_DECL
    for (int s = 0; s < p.T; ++s) {
      // K-chunk kc of the A operand is the h slice of CTA kc.  The MMAs of a chunk are issued as soon as THAT chunk has
      // landed (own chunk first), so the tensor core works while the peers' copies are still in flight.
      if (s > 0 && elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
          if (kc != jt) mbar_expect_tx(hready + 8 * kc, (uint32_t)p.rows * 128u);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kc = (jt + i) & 3;
        if (s > 0) { LT(lt_a, mbar_wait(hready + 8 * kc, (uint32_t)((s - 1) & 1))) }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t ad = umma_desc<128>(abuf + kc * (128 * 128)), bd = umma_desc<128>(wbuf + kc * (256 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_acc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(tfull);                       // accumulators of this step ready (own epilogue)
        umma_commit_multicast(afree, 0xF);        // ... and this CTA no longer reads h_{s-1} (tell all four CTAs)
      }
      __syncwarp();
    }
#ifdef VTD_TIMERS
    if (lane == 0 && p.timers && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { p.timers[0] = clock64() - lt_t0; p.timers[1] = lt_a; }
#endif
  } else {
    // ---- epilogue: gates, cell update, h exchange
    const int q = warp & 3, half = (warp - 2) >> 2;   // half = which L_UPT-unit part of the 64 units
    const int m = q * 32 + lane;                        // sequence row inside the tile
    const int b = b0 + m;
    const bool valid = b < B;
    const bool active = q * 32 < p.rows;                // with 64 rows per cluster the upper two lane quarters carry no sequence
    const uint32_t epi_threads = (uint32_t)(p.rows * L_PARTS);
    const bool issuer = q == 0 && half == 0 && lane == 0;   // one thread of an always-active warp pushes the slice to the peers
    const int bb = valid ? b : 0;
    float c[L_UPT];
#pragma unroll
    for (int i = 0; i < L_UPT; ++i) c[i] = 0.f;
    // remote addresses of this thread's 4 x 16-byte h chunks in every CTA of the cluster
    const uint32_t own_chunk = abuf + jt * (128 * 128);             // K-chunk jt of the A operand (local address)
    const uint32_t own_row = own_chunk + m * 128;
    uint32_t dst_chunk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) dst_chunk[r] = map_to_cta(own_chunk, (uint32_t)r);
    uint32_t hrdy[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) hrdy[r] = map_to_cta(hready + 8 * jt, (uint32_t)r);   // "chunk jt has landed" in CTA r
    LT_DECL
    for (int s = 0; active && s < p.T; ++s) {
      const int t = dir == 0 ? s : p.T - 1 - s;
      const bf16* __restrict__ xp = p.xproj + (((size_t)bb * p.T + t) * 2 + dir) * 1024 + jt * 256 + half * L_UPT;
      bf16* __restrict__ so = p.seq_out + ((size_t)bb * p.T + t) * 512 + dir * 256 + jt * 64 + half * L_UPT;
      // input projection of this step: 4 gates x 32 units bf16 = 16 x 16 bytes, requested BEFORE the accumulator wait so
      // the L2/DRAM latency overlaps the MMAs (loading them chunk by chunk inside the gate loop cost 8 exposed round
      // trips per step: long_scoreboard was 49 % of the stalls)
#ifdef VTD_TIMERS
      const long long lt_x0 = clock64();
#endif
      uint4 xr[4][L_CH];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < L_CH; ++j)
          xr[g][j] = valid ? __ldg(reinterpret_cast<const uint4*>(xp + g * 64) + j) : make_uint4(0u, 0u, 0u, 0u);
      // xproj of one layer is ~100 MB: the rows of the steps after this one are pulled into L2 now, so that the loads
      // above find them there (DRAM latency is longer than the MMAs of a step; long_scoreboard was the top stall)
      if (valid && s + 2 < p.T) {
        const int t2 = dir == 0 ? s + 2 : p.T - 3 - s;
        const bf16* xn = p.xproj + (((size_t)bb * p.T + t2) * 2 + dir) * 1024 + jt * 256 + half * L_UPT;
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + g * 64));
      }
#ifdef VTD_TIMERS
      lt_d += clock64() - lt_x0;
#endif
      LT(lt_a, mbar_wait(tfull, (uint32_t)(s & 1)))
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef VTD_TIMERS
      const long long lt_m0 = clock64();
#endif
      uint4 hw[L_CH];
#pragma unroll
      for (int j = 0; j < L_CH; ++j) {                  // 8 hidden units at a time
        uint32_t vi[8], vf[8], vg[8], vo[8];
        const uint32_t tb = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * L_UPT + j * 8);
        tmem_ld8(tb, vi); tmem_ld8(tb + 64, vf); tmem_ld8(tb + 128, vg); tmem_ld8(tb + 192, vo);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bf16x2* xi2 = reinterpret_cast<const bf16x2*>(&xr[0][j]);
        const bf16x2* xf2 = reinterpret_cast<const bf16x2*>(&xr[1][j]);
        const bf16x2* xg2 = reinterpret_cast<const bf16x2*>(&xr[2][j]);
        const bf16x2* xo2 = reinterpret_cast<const bf16x2*>(&xr[3][j]);
        float hv[8];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 xi = unpack2(xi2[e2]), xf = unpack2(xf2[e2]);
          const float2 xg = unpack2(xg2[e2]), xo = unpack2(xo2[e2]);
          const float xiv[2] = {xi.x, xi.y}, xfv[2] = {xf.x, xf.y}, xgv[2] = {xg.x, xg.y}, xov[2] = {xo.x, xo.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int u = e2 * 2 + e;
            const float gi = __uint_as_float(vi[u]) + xiv[e], gf = __uint_as_float(vf[u]) + xfv[e];
            const float gg = __uint_as_float(vg[u]) + xgv[e], go = __uint_as_float(vo[u]) + xov[e];
            const float cn = sigmoid_approx(gf) * c[j * 8 + u] + sigmoid_approx(gi) * tanh_approx(gg);
            c[j * 8 + u] = cn;
            hv[u] = valid ? sigmoid_approx(go) * tanh_approx(cn) : 0.f;
          }
        }
        bf16x2 a0 = pack2(hv[0], hv[1]), a1 = pack2(hv[2], hv[3]);
        bf16x2 a2 = pack2(hv[4], hv[5]), a3 = pack2(hv[6], hv[7]);
        uint4 u4;
        u4.x = *reinterpret_cast<uint32_t*>(&a0); u4.y = *reinterpret_cast<uint32_t*>(&a1);
        u4.z = *reinterpret_cast<uint32_t*>(&a2); u4.w = *reinterpret_cast<uint32_t*>(&a3);
        hw[j] = u4;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#ifdef VTD_TIMERS
      lt_c += clock64() - lt_m0;
#endif
      if (s + 1 < p.T) {
        LT(lt_b, mbar_wait(afree, (uint32_t)(s & 1)))            // nobody's MMAs still read h_{s-1}
        // h_s slice -> A operand of all 4 CTAs: 16-byte chunk j of the row goes to physical chunk j ^ (row & 7)
#ifdef VTD_TIMERS
        const long long lt_s0 = clock64();
#endif
        // h_s slice: 16-byte chunk c of row m goes to physical chunk c ^ (m & 7) of K-chunk jt of the OWN A buffer; that
        // K-chunk (128 rows x 128 B = 16 KB, contiguous) is then pushed to the three peers with one bulk DSMEM copy each,
        // whose bytes complete the peers' hready barriers (16-byte st.async packets moved the same bytes ~3x slower)
#pragma unroll
        for (int j = 0; j < L_CH; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(own_row + ((uint32_t)((half * L_CH + j) ^ (m & 7)) << 4)),
                       "r"(hw[j].x), "r"(hw[j].y), "r"(hw[j].z), "r"(hw[j].w) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");      // all active epilogue warps have written their part
        if (issuer) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r != jt)
              asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(dst_chunk[r]), "r"(own_chunk), "r"((uint32_t)p.rows * 128u), "r"(hrdy[r]) : "memory");
          mbar_arrive(hready + 8 * jt);                     // own slice is in place (release; the MMA thread's wait acquires)
        }
#ifdef VTD_TIMERS
        lt_e += clock64() - lt_s0;
#endif
      }
#ifdef VTD_TIMERS
      const long long lt_g0 = clock64();
#endif
      if (valid) {                                      // layer output (global) after the exchange: off the critical path
        uint4* sp = reinterpret_cast<uint4*>(so);
#pragma unroll
        for (int j = 0; j < L_CH; ++j) sp[j] = hw[j];
      }
#ifdef VTD_TIMERS
      lt_f += clock64() - lt_g0;
#endif
    }
#ifdef VTD_TIMERS
    if (issuer && p.timers && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      p.timers[2] = clock64() - lt_t0; p.timers[3] = lt_a; p.timers[4] = lt_b; p.timers[5] = lt_c; p.timers[6] = lt_d; p.timers[7] = lt_e; p.timers[8] = lt_f;
    }
#endif
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(256u) : "memory");
  }
  cluster_arrive();
  cluster_wait();                                       // no CTA exits while a peer may still address its shared memory
}

}  // namespace

struct LstmPlan {
  CUtensorMap map_whh;
};

LstmPlan* lstm_plan_create(const void* whh /*[2*1024][256] bf16, rows (dir, unit tile, gate, unit)*/, std::string* err) {
  tc::EncodeTiledFn enc = tc::get_encode();
  if (!enc) { if (err) *err = "cuTensorMapEncodeTiled not available from the driver"; return nullptr; }
  LstmPlan* pl = new LstmPlan();
  cuuint64_t dims[2] = {256, 2048};
  cuuint64_t strides[1] = {256 * 2};
  cuuint32_t box[2] = {64, 256};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&pl->map_whh, VTD_TMAP_16, 2, const_cast<void*>(whh), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(W_hh) failed: " + std::to_string((int)r);
    delete pl;
    return nullptr;
  }
  return pl;
}

void lstm_plan_destroy(LstmPlan* p) { delete p; }

cudaError_t bilstm_layer_tcgen05(const LstmPlan* pl, const void* xproj /*bf16*/, void* seq_out, int B, int T, cudaStream_t s,
                                 LaunchCounter* lc, const int* n_dyn, int n_first) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  static PerDeviceFlag attr_done;
  {
    cudaError_t e = once_per_device(attr_done, [] {
      return cudaFuncSetAttribute(bilstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM);
    });
    if (e != cudaSuccess) return e;
  }
  LstmParams p;
  p.xproj = reinterpret_cast<const bf16*>(xproj); p.seq_out = reinterpret_cast<bf16*>(seq_out); p.B = B; p.T = T;
  p.timers = nullptr;
  p.n_dyn = n_dyn; p.n_first = n_first;
  // 64 sequences per cluster when all clusters are still resident at once: the MMA costs the same (M = 128 either way),
  // the per-step h exchange (DSMEM, ~9 B/cycle/SM measured) and the gate math per CTA halve
  p.rows = (((B + 63) / 64) * 8 <= tc::sm_count() && !dev_env("VTD_LSTM_ROWS128")) ? 64 : 128;
  dim3 grid(4, 2, (B + p.rows - 1) / p.rows);
#ifdef VTD_TIMERS
  if (getenv("VTD_TIMERS")) {
    static long long* tb = nullptr;
    if (!tb) cudaMalloc(&tb, 16 * sizeof(long long));
    p.timers = tb;
    bilstm_persistent_kernel<<<grid, L_THREADS, L_SMEM, s>>>(pl->map_whh, p);
    cudaStreamSynchronize(s);
    long long h[9]; cudaMemcpy(h, tb, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "LSTM B=%d T=%d | mma tot %lld wait_hready %lld | epi tot %lld wait_tfull %lld wait_afree %lld math %lld xload %lld stasync %lld gstore %lld | per step %lld\n",
            B, T, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[2] / T);
    if (lc) lc->n++;
    return cudaGetLastError();
  }
#endif
  bilstm_persistent_kernel<<<grid, L_THREADS, L_SMEM, s>>>(pl->map_whh, p);
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
