// cta_group::2 check: a cluster of two CTAs computes D[256 x 128] = A[256 x 64] * B[128 x 64]^T with ONE stream of
// tcgen05.mma.cta_group::2 issued by the leader: CTA r holds rows 128r.. of A and rows 64r.. of B (N split in halves) in
// its own shared memory at the SAME offsets, and gets rows 128r.. of D (all 128 columns) in its own TMEM.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* a /*[256][64]*/, const __nv_bfloat16* b /*[128][64]*/, float* out /*[256][128]*/) {
  __shared__ __align__(1024) uint8_t sa[128 * 128];
  __shared__ __align__(1024) uint8_t sb[64 * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int R = i / 64, e = i % 64, c = e / 8;
    reinterpret_cast<__nv_bfloat16*>(sa)[R * 64 + ((c ^ (R & 7)) * 8) + (e % 8)] = a[(rank * 128 + R) * 64 + e];
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int R = i / 64, e = i % 64, c = e / 8;
    reinterpret_cast<__nv_bfloat16*>(sb)[R * 64 + ((c ^ (R & 7)) * 8) + (e % 8)] = b[(rank * 64 + R) * 64 + e];
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // both CTAs' operands and barriers are ready
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t ad = umma_desc<128>(smem_u32(sa)) + (uint64_t)(ks * 2);
      const uint64_t bd = umma_desc<128>(smem_u32(sb)) + (uint64_t)(ks * 2);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  for (int h = 0; h < 4; ++h) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(rank * 128 + warp * 32 + lane) * 128 + h * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> a(256 * 64), b(128 * 64);
  std::vector<float> af(a.size()), bf(b.size());
  for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)((int)(i * 37 % 29) - 14) / 8.f; a[i] = __float2bfloat16(af[i]); }
  for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)((int)(i * 13 % 17) - 8) / 4.f; b[i] = __float2bfloat16(bf[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 256 * 128 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 256 * 128 * 4);
  k<<<2, 128>>>(da, db, dout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> out(256 * 128);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int kk = 0; kk < 64; ++kk) s += (double)af[m * 64 + kk] * (double)bf[n * 64 + kk];
      const double err = fabs(s - out[m * 128 + n]);
      if (!(err < 1e-3)) ++bad;
      if (err == err) maxerr = fmax(maxerr, err);
    }
  printf("umma_2cta: %s, max |err| = %g, bad = %d of %d\n", cudaGetErrorString(e), maxerr, bad, 256 * 128);
  return 0;
}
