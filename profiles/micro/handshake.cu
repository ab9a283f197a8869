// Micro-benchmark of the producer/consumer mbarrier ring used by conv_tcgen05.cu (no data, no MMAs): cycles per k-step
// of the bare handshake under different variants.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handshake handshake.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

__device__ __forceinline__ void wait_simple(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

// variant bits: 1 = consumer releases with mbarrier.arrive instead of tcgen05.commit; 2 = only lane 0 waits (then syncwarp)
//               4 = simple wait loop without the clock64 watchdog; 8 = no tcgen05.fence in the consumer
template <int V>
__global__ void __launch_bounds__(64, 1) ring(int stages, int iters, long long* out) {
  __shared__ uint64_t bars[64];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
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
    for (int i = 0; i < iters; ++i) {
      if (V & 2) { if (lane == 0) { if (V & 4) wait_simple(empty0 + 8 * stage, phase ^ 1); else mbar_wait(empty0 + 8 * stage, phase ^ 1); } __syncwarp(); }
      else { if (V & 4) wait_simple(empty0 + 8 * stage, phase ^ 1); else mbar_wait(empty0 + 8 * stage, phase ^ 1); }
      if (elect_one()) mbar_arrive(full0 + 8 * stage);
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (V & 2) { if (lane == 0) { if (V & 4) wait_simple(full0 + 8 * stage, phase); else mbar_wait(full0 + 8 * stage, phase); } __syncwarp(); }
      else { if (V & 4) wait_simple(full0 + 8 * stage, phase); else mbar_wait(full0 + 8 * stage, phase); }
      if (!(V & 8)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        if (V & 1) mbar_arrive(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 2 + warp] = t1 - t0;
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
}

template <int V> void run(const char* name, int stages, int iters) {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  ring<V><<<148, 64>>>(stages, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 296; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-40s stages=%2d  %.1f cycles/k-step  (%s)\n", name, stages, (double)mx / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int it = 20000;
  for (int st : {2, 4, 9}) {
    run<0>("commit, all lanes wait, watchdog wait", st, it);
    run<1>("arrive, all lanes wait, watchdog wait", st, it);
    run<2>("commit, lane0 waits", st, it);
    run<4>("commit, all lanes, simple wait", st, it);
    run<5>("arrive, all lanes, simple wait", st, it);
    run<6>("commit, lane0, simple wait", st, it);
    run<7>("arrive, lane0, simple wait", st, it);
    run<14>("commit, lane0, simple wait, no fence", st, it);
  }
  return 0;
}
