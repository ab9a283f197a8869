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
// and the layer output (global) -> two cluster barriers order "all MMAs have read h_{t-1}" before "h_t is written"
// before "next MMAs".  Latency per step is a few microseconds instead of one kernel launch per step.
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
  const float* xproj;     // [B][T][2][1024]
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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x;                            // == rank in cluster: hidden units 64*jt ..
  const int dir = blockIdx.y;
  const int b0 = blockIdx.z * 128;

  if (warp == 0 && lane == 0) {
    mbar_init(wfull, 1);
    mbar_init(tfull, 1);
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
    for (int s = 0; s < p.T; ++s) { cluster_arrive(); cluster_wait(); cluster_arrive(); cluster_wait(); }
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = umma_idesc(256);
    mbar_wait(wfull, 0);
    for (int s = 0; s < p.T; ++s) {
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t ad = umma_desc<128>(abuf + kc * (128 * 128)), bd = umma_desc<128>(wbuf + kc * (256 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_acc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
        }
        umma_commit(tfull);
      }
      __syncwarp();
      cluster_arrive(); cluster_wait();                 // #1: every CTA's MMAs of this step have completed
      cluster_arrive(); cluster_wait();                 // #2: h_t has landed in every CTA
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
    for (int s = 0; s < p.T; ++s) {
      const int t = dir == 0 ? s : p.T - 1 - s;
      const float* __restrict__ xp = p.xproj + (((size_t)bb * p.T + t) * 2 + dir) * 1024 + jt * 256 + half * 32;
      bf16* __restrict__ so = p.seq_out + ((size_t)bb * p.T + t) * 512 + dir * 256 + jt * 64 + half * 32;
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + g * 64));
      }
      mbar_wait(tfull, (uint32_t)(s & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      cluster_arrive();                                 // #1 (own MMAs are complete)
      uint4 hw[4];
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {                  // 16 hidden units at a time
        uint32_t vi[16], vf[16], vg[16], vo[16];
        const uint32_t tb = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32 + qq * 16);
        tmem_ld16(tb, vi); tmem_ld16(tb + 64, vf); tmem_ld16(tb + 128, vg); tmem_ld16(tb + 192, vo);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float hv[16];
#pragma unroll
        for (int u4 = 0; u4 < 16; u4 += 4) {
          const int u = qq * 16 + u4;
          float4 xi = make_float4(0.f, 0.f, 0.f, 0.f), xf = xi, xg = xi, xo = xi;
          if (valid) {
            xi = __ldg(reinterpret_cast<const float4*>(xp + u));
            xf = __ldg(reinterpret_cast<const float4*>(xp + 64 + u));
            xg = __ldg(reinterpret_cast<const float4*>(xp + 128 + u));
            xo = __ldg(reinterpret_cast<const float4*>(xp + 192 + u));
          }
          const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
          const float xgv[4] = {xg.x, xg.y, xg.z, xg.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float gi = __uint_as_float(vi[u4 + e]) + xiv[e], gf = __uint_as_float(vf[u4 + e]) + xfv[e];
            const float gg = __uint_as_float(vg[u4 + e]) + xgv[e], go = __uint_as_float(vo[u4 + e]) + xov[e];
            const float cn = fast_sigmoid(gf) * c[u + e] + fast_sigmoid(gi) * fast_tanh(gg);
            c[u + e] = cn;
            hv[u4 + e] = valid ? fast_sigmoid(go) * fast_tanh(cn) : 0.f;
          }
        }
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2) {
          __nv_bfloat162 a0 = __floats2bfloat162_rn(hv[8 * w2 + 0], hv[8 * w2 + 1]);
          __nv_bfloat162 a1 = __floats2bfloat162_rn(hv[8 * w2 + 2], hv[8 * w2 + 3]);
          __nv_bfloat162 a2 = __floats2bfloat162_rn(hv[8 * w2 + 4], hv[8 * w2 + 5]);
          __nv_bfloat162 a3 = __floats2bfloat162_rn(hv[8 * w2 + 6], hv[8 * w2 + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&a0); u.y = *reinterpret_cast<uint32_t*>(&a1);
          u.z = *reinterpret_cast<uint32_t*>(&a2); u.w = *reinterpret_cast<uint32_t*>(&a3);
          hw[qq * 2 + w2] = u;
        }
      }
      if (valid) {
        uint4* sp = reinterpret_cast<uint4*>(so);
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[j] = hw[j];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      cluster_wait();                                   // #1: nobody's MMAs still read h_{t-1}
      // h_t slice -> A operand of all 4 CTAs: 16-byte chunk j of the row goes to physical chunk j ^ (row & 7)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) st_cluster_v4(dst[r] + ((uint32_t)((half * 4 + j) ^ (m & 7)) << 4), hw[j]);
      asm volatile("fence.proxy.async;" ::: "memory");
      cluster_arrive(); cluster_wait();                 // #2
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

cudaError_t bilstm_layer_tcgen05(const LstmPlan* pl, const float* xproj, void* seq_out, int B, int T, cudaStream_t s,
                                 LaunchCounter* lc) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bilstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  LstmParams p;
  p.xproj = xproj; p.seq_out = reinterpret_cast<bf16*>(seq_out); p.B = B; p.T = T;
  dim3 grid(4, 2, (B + 127) / 128);
  bilstm_persistent_kernel<<<grid, L_THREADS, L_SMEM, s>>>(pl->map_whh, p);
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
