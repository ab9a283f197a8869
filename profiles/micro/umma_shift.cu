// Can a SWIZZLE_128B K-major A descriptor start at a row that is NOT a multiple of 8 (start address not 1024-aligned) with
// 8-row groups SBO apart -- i.e. can a 3x3 convolution read its shifted taps straight out of ONE halo patch in shared
// memory?  Rows are stored as TMA stores them: 16-byte chunk c of absolute row R at chunk c ^ (R & 7).
// Tries the descriptor's base-offset field = 0 and = (start >> 7) & 7.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../video_text_detection_system_b200/csrc/tc_common.cuh"
using namespace vtd::tc;

constexpr int NROWS = 208, PW = 10;

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* a_rows /*[NROWS][64]*/, const __nv_bfloat16* b_nk /*[64][64]*/,
                                            int r0, int use_base_offset, float* out) {
  __shared__ __align__(1024) uint8_t sa[NROWS * 128];
  __shared__ __align__(1024) uint8_t sb[64 * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NROWS * 64; i += 128) {
    const int R = i / 64, e = i % 64, c = e / 8;
    reinterpret_cast<__nv_bfloat16*>(sa)[R * 64 + ((c ^ (R & 7)) * 8) + (e % 8)] = a_rows[i];
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int R = i / 64, e = i % 64, c = e / 8;
    reinterpret_cast<__nv_bfloat16*>(sb)[R * 64 + ((c ^ (R & 7)) * 8) + (e % 8)] = b_nk[i];
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
    const uint32_t start = smem_u32(sa) + r0 * 128;
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t ad = 0;
      ad |= (uint64_t)(((start + ks * 32) & 0x3FFFF) >> 4);
      ad |= (uint64_t)1 << 16;
      ad |= (uint64_t)((PW * 128) >> 4) << 32;          // 8-row groups PW rows apart
      ad |= (uint64_t)1 << 46;
      if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7) << 49;
      ad |= (uint64_t)2 << 61;
      const uint64_t bd = umma_desc<128>(smem_u32(sb)) + (uint64_t)(ks * 2);
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
  std::vector<__nv_bfloat16> a(NROWS * 64), b(64 * 64);
  std::vector<float> af(a.size()), bf(b.size());
  for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)((int)(i * 37 % 29) - 14) / 8.f; a[i] = __float2bfloat16(af[i]); }
  for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)((int)(i * 13 % 17) - 8) / 4.f; b[i] = __float2bfloat16(bf[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  for (int r0 : {0, 8, 1, 11, 21}) for (int ubo = 0; ubo < 2; ++ubo) {
    k<<<1, 128>>>(da, db, r0, ubo, dout);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 64);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        const int R = r0 + (m / 8) * PW + (m % 8);
        double s = 0;
        for (int kk = 0; kk < 64; ++kk) s += (double)af[R * 64 + kk] * (double)bf[n * 64 + kk];
        maxerr = fmax(maxerr, fabs(s - out[m * 64 + n]));
      }
    printf("umma_shift: start row %2d base_offset field %s: %s, max |err| = %g\n", r0, ubo ? "set" : "0  ", cudaGetErrorString(e), maxerr);
  }
  return 0;
}
