// Handshake ring (as handshake.cu) plus 8 "epilogue" warps that wait on a per-tile barrier the consumer commits every
// `per_tile` slots -- how much do spinning waiters slow the producer/consumer handshake down?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

__device__ __forceinline__ bool try_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return done != 0;
}

// EPI: 0 = no epilogue warps wait (they exit), 1 = all lanes spin (mbar_wait), 2 = lane 0 spins + syncwarp,
//      3 = all lanes, try_wait with a 10 us suspend hint, 4 = all lanes with __nanosleep(64) backoff, 5 = lane 0 + hint
template <int EPI>
__global__ void __launch_bounds__(320, 1) ring(int stages, int tiles, int per_tile, long long* out) {
  __shared__ uint64_t bars[64];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * 16, tfull0 = full0 + 8 * 32, tempty0 = full0 + 8 * 36;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, EPI ? 8 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < tiles * per_tile; ++i) {
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      if (elect_one()) mbar_arrive(full0 + 8 * stage);
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int t = 0; t < tiles; ++t) {
      if (EPI) mbar_wait(tempty0 + 8 * as, aphase ^ 1);
      for (int i = 0; i < per_tile; ++i) {
        mbar_wait(full0 + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) { umma_commit(empty0 + 8 * stage); if (i == per_tile - 1) umma_commit(tfull0 + 8 * as); }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++as == 4) { as = 0; aphase ^= 1; }
    }
  } else if (EPI) {
    int as = 0; uint32_t aphase = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t b = tfull0 + 8 * as;
      if (EPI == 1) mbar_wait(b, aphase);
      else if (EPI == 2) { if (lane == 0) mbar_wait(b, aphase); __syncwarp(); }
      else if (EPI == 3) { while (!try_hint(b, aphase, 10000)) {} }
      else if (EPI == 4) { while (!mbar_try(b, aphase)) __nanosleep(64); }
      else if (EPI == 5) { if (lane == 0) { while (!try_hint(b, aphase, 10000)) {} } __syncwarp(); }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == 4) { as = 0; aphase ^= 1; }
    }
  }
  long long t1 = clock64();
  if (lane == 0 && warp < 2) out[blockIdx.x * 2 + warp] = t1 - t0;
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
}

template <int EPI> void run(const char* name, int stages, int tiles, int per_tile) {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaMemset(d, 0, 148 * 2 * sizeof(long long));
  ring<EPI><<<148, 320>>>(stages, tiles, per_tile, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 296; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-44s slots/tile=%2d  %.1f cycles/slot  %.0f cycles/tile (%s)\n", name, per_tile, (double)mx / (tiles * per_tile),
         (double)mx / tiles, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int pt : {1, 3, 9, 36}) {
    const int tiles = 36000 / pt;
    run<0>("no epilogue waiters", 4, tiles, pt);
    run<1>("8 warps, all lanes spin", 4, tiles, pt);
    run<2>("8 warps, lane 0 spins", 4, tiles, pt);
    run<3>("8 warps, all lanes, suspend hint", 4, tiles, pt);
    run<4>("8 warps, all lanes, nanosleep(64)", 4, tiles, pt);
    run<5>("8 warps, lane 0, suspend hint", 4, tiles, pt);
  }
  return 0;
}
