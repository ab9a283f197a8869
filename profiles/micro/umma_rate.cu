// How fast does tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) issue back to back when its SWIZZLE_128B K-major A operand
// starts at a row that is not a multiple of 8 (a shifted 3x3 tap read straight out of a halo patch) and its 8-row groups
// are PW rows apart?  Prints cycles per MMA for N = 64/128/256, several start rows and group pitches.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../video_text_detection_system_b200/csrc/common.cuh"
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

constexpr int NROWS = 16 * 20 + 64;

__global__ void __launch_bounds__(128, 1) k(int N, int r0, int PW, int reps, int ksteps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                       // NROWS x 128 B
  uint8_t* sb = smem + ((NROWS * 128 + 1023) & ~1023);   // 256 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (NROWS * 128 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(N);
    const uint32_t start = smem_u32(sa) + r0 * 128;
    uint64_t adv[4], bdv[4];
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t ad = 0;
      ad |= (uint64_t)(((start + ks * 32) & 0x3FFFF) >> 4);
      ad |= (uint64_t)1 << 16;
      ad |= (uint64_t)((PW * 128) >> 4) << 32;
      ad |= (uint64_t)1 << 46;
      ad |= (uint64_t)((start >> 7) & 7) << 49;
      ad |= (uint64_t)2 << 61;
      adv[ks] = ad;
      bdv[ks] = umma_desc<128>(smem_u32(sb)) + (uint64_t)(ks * 2);
    }
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16(tm, adv[ks], bdv[ks], idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    *cycles = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = ((NROWS * 128 + 1023) & ~1023) + 256 * 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 256, ksteps = 4;
  for (int N : {64, 128, 256})
    for (int PW : {8, 10, 18})
      for (int r0 : {0, 1, 2, 8, 10, 11, 19}) {
        long long c = 0;
        for (int t = 0; t < 2; ++t) { k<<<1, 128, smem>>>(N, r0, PW, reps, ksteps, d); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError();
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d pitch=%2d rows, start row %2d: %6.1f cycles/MMA  %s\n", N, PW, r0, (double)c / (reps * ksteps), e ? cudaGetErrorString(e) : "");
      }
  return 0;
}
