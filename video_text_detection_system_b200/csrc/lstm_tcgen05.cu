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

namespace vtd {
namespace {

using namespace tc;

constexpr int L_THREADS = 320;                 // warp 0: weight loader, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int L_W_BYTES = 4 * 256 * 128;       // 4 K-chunks x 256 rows x 128 B
constexpr int L_A_BYTES = 4 * 128 * 128;       // 4 K-chunks x 128 rows x 128 B
constexpr int L_SMEM = L_W_BYTES + L_A_BYTES + 1024 + 64;

struct LstmParams {
  const bf16* xproj;      // [B][T][2][1024] bf16
  bf16* seq_out;          // [B][T][512]
  int B, T;
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
  const uint32_t hready = wfull + 24;     // 1 arrival (own MMA thread, expect_tx 64 KB) + the st.async bytes of h_t from all 4 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x;                            // == rank in cluster: hidden units 64*jt ..
  const int dir = blockIdx.y;
  const int b0 = blockIdx.z * 128;

  if (warp == 0 && lane == 0) {
    mbar_init(wfull, 1);
    mbar_init(tfull, 1);
    mbar_init(afree, 4);
    mbar_init(hready, 1);
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
    for (int s = 0; s < p.T; ++s) {
      if (s > 0) {                                      // h_{s-1} of all 256 units (64 KB, written by st.async) has landed here
        if (elect_one()) mbar_expect_tx(hready, L_A_BYTES);
        __syncwarp();
        mbar_wait(hready, (uint32_t)((s - 1) & 1));
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t ad = umma_desc<128>(abuf + kc * (128 * 128)), bd = umma_desc<128>(wbuf + kc * (256 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_acc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
        }
        umma_commit(tfull);                       // accumulators of this step ready (own epilogue)
        umma_commit_multicast(afree, 0xF);        // ... and this CTA no longer reads h_{s-1} (tell all four CTAs)
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: gates, cell update, h exchange
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;                        // sequence row inside the tile
    const int b = b0 + m;
    const bool valid = b < p.B;
    const int bb = valid ? b : 0;
    float c[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = 0.f;
    // remote addresses of this thread's 4 x 16-byte h chunks in every CTA of the cluster
    uint32_t dst[4];
    {
      const uint32_t row = abuf + jt * (128 * 128) + m * 128;      // K-chunk jt, row m (local address)
#pragma unroll
      for (int r = 0; r < 4; ++r) dst[r] = map_to_cta(row, (uint32_t)r);
    }
    uint32_t hrdy[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) hrdy[r] = map_to_cta(hready, (uint32_t)r);
    for (int s = 0; s < p.T; ++s) {
      const int t = dir == 0 ? s : p.T - 1 - s;
      const bf16* __restrict__ xp = p.xproj + (((size_t)bb * p.T + t) * 2 + dir) * 1024 + jt * 256 + half * 32;
      bf16* __restrict__ so = p.seq_out + ((size_t)bb * p.T + t) * 512 + dir * 256 + jt * 64 + half * 32;
      // input projection of this step: 4 gates x 32 units bf16 = 16 x 16 bytes, requested BEFORE the accumulator wait so
      // the L2/DRAM latency overlaps the MMAs (loading them chunk by chunk inside the gate loop cost 8 exposed round
      // trips per step: long_scoreboard was 49 % of the stalls)
      uint4 xr[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          xr[g][j] = valid ? __ldg(reinterpret_cast<const uint4*>(xp + g * 64) + j) : make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(tfull, (uint32_t)(s & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint4 hw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {                     // 8 hidden units at a time
        uint32_t vi[8], vf[8], vg[8], vo[8];
        const uint32_t tb = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32 + j * 8);
        tmem_ld8(tb, vi); tmem_ld8(tb + 64, vf); tmem_ld8(tb + 128, vg); tmem_ld8(tb + 192, vo);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const __nv_bfloat162* xi2 = reinterpret_cast<const __nv_bfloat162*>(&xr[0][j]);
        const __nv_bfloat162* xf2 = reinterpret_cast<const __nv_bfloat162*>(&xr[1][j]);
        const __nv_bfloat162* xg2 = reinterpret_cast<const __nv_bfloat162*>(&xr[2][j]);
        const __nv_bfloat162* xo2 = reinterpret_cast<const __nv_bfloat162*>(&xr[3][j]);
        float hv[8];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 xi = __bfloat1622float2(xi2[e2]), xf = __bfloat1622float2(xf2[e2]);
          const float2 xg = __bfloat1622float2(xg2[e2]), xo = __bfloat1622float2(xo2[e2]);
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
        __nv_bfloat162 a0 = __floats2bfloat162_rn(hv[0], hv[1]), a1 = __floats2bfloat162_rn(hv[2], hv[3]);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(hv[4], hv[5]), a3 = __floats2bfloat162_rn(hv[6], hv[7]);
        uint4 u4;
        u4.x = *reinterpret_cast<uint32_t*>(&a0); u4.y = *reinterpret_cast<uint32_t*>(&a1);
        u4.z = *reinterpret_cast<uint32_t*>(&a2); u4.w = *reinterpret_cast<uint32_t*>(&a3);
        hw[j] = u4;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (s + 1 < p.T) {
        mbar_wait(afree, (uint32_t)(s & 1));            // nobody's MMAs still read h_{s-1}
        // h_s slice -> A operand of all 4 CTAs: 16-byte chunk j of the row goes to physical chunk j ^ (row & 7)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j) st_async_v4(dst[r] + ((uint32_t)((half * 4 + j) ^ (m & 7)) << 4), hw[j], hrdy[r]);
      }
      if (valid) {                                      // layer output (global) after the exchange: off the critical path
        uint4* sp = reinterpret_cast<uint4*>(so);
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[j] = hw[j];
      }
    }
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
  CUresult r = enc(&pl->map_whh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(whh), dims, strides, box, es,
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
                                 LaunchCounter* lc) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bilstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  LstmParams p;
  p.xproj = reinterpret_cast<const bf16*>(xproj); p.seq_out = reinterpret_cast<bf16*>(seq_out); p.B = B; p.T = T;
  dim3 grid(4, 2, (B + 127) / 128);
  bilstm_persistent_kernel<<<grid, L_THREADS, L_SMEM, s>>>(pl->map_whh, p);
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
