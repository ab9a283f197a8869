// Does tcgen05.mma accept a NO-SWIZZLE K-major A descriptor whose rows OVERLAP in shared memory (row pitch 16 bytes =
// the core-matrix row pitch, K-direction core-matrix stride LBO = 16 bytes)?  That layout is exactly the 7x7 stride-2
// stem's im2col: output pixel x reads the 32 bf16 starting at input element 8x.  D[128x64] = A[128x32] * B[64x32]^T.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                   // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* a_flat, const __nv_bfloat16* b_nk, float* out) {
  __shared__ __align__(1024) uint8_t sa[8192];
  __shared__ __align__(1024) uint8_t sb[64 * 32 * 2];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: flat copy of (128*8 + 32) elements
  for (int i = threadIdx.x; i < 128 * 8 + 32; i += 128) reinterpret_cast<__nv_bfloat16*>(sa)[i] = a_flat[i];
  // B canonical no-swizzle: [k/8][n/8][n%8][k%8]
  for (int i = threadIdx.x; i < 64 * 32; i += 128) {
    const int n = i / 32, kk = i % 32;
    reinterpret_cast<__nv_bfloat16*>(sb)[((kk / 8) * 8 + n / 8) * 64 + (n % 8) * 8 + (kk % 8)] = b_nk[i];
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(64);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t ad = desc_nosw(smem_u32(sa) + ks * 32, 16, 128);          // rows 16 B apart, K core matrices 16 B apart
      const uint64_t bd = desc_nosw(smem_u32(sb) + ks * 2 * 1024, 1024, 128);  // 2 K core matrices per K=16 step
      umma_f16(tm, ad, bd, idesc, ks ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + h * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> a(128 * 8 + 32), b(64 * 32);
  std::vector<float> af(a.size()), bf(b.size());
  for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)((int)(i * 37 % 23) - 11) / 8.f; a[i] = __float2bfloat16(af[i]); }
  for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)((int)(i * 13 % 17) - 8) / 4.f; b[i] = __float2bfloat16(bf[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  k<<<1, 128>>>(da, db, dout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> out(128 * 64);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int kk = 0; kk < 32; ++kk) s += (double)af[m * 8 + kk] * (double)bf[n * 32 + kk];
      maxerr = fmax(maxerr, fabs(s - out[m * 64 + n]));
    }
  printf("umma_nosw: %s, max |err| = %g (expect 0: all products exact)\n", cudaGetErrorString(e), maxerr);
  return 0;
}
