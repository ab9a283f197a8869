// Stage 4b (last step): softmax + greedy decode with the REFERENCE's semantics.
//
// Replaces torch.softmax(outputs, dim=2) (text_recognizer.py:126) and TextRecognizer._decode_prediction
// (:142-167).  That decode is not canonical CTC:
//   * blanks (id 0) are skipped WITHOUT resetting prev_char   ([a,0,a,b] -> "ab");
//   * an id equal to prev_char is skipped;
//   * <unk> (id V-1 = 96) is dropped from the text but still becomes prev_char;
//   * the confidence of the k-th emitted character is max(prediction[k-1]) -- the row is indexed by the
//     number of characters emitted so far, not by the timestep; the result is their mean (0.0 if none).
// `canonical != 0` switches to textbook CTC collapse (blank resets prev), for users who want it.
//
// One CTA per sequence, one WARP per timestep (the rows of a sequence are independent until the collapse): the 32
// lanes stride over the V classes, max / sum-of-exp, shuffle argmax with lowest-index tie-break (torch.argmax); after
// a CTA barrier thread 0 runs the <=64-step collapse.  (One warp walking the T rows of a sequence one after the other
// was a chain of ~T x 3 dependent L2 round trips: 73 us for 800 x 31 rows; the bytes are 12 MB.)
// HBM-bound: reads B*T*V*4 bytes once, writes ~T bytes.
#include "common.cuh"
#include "../../include/vtd.h"

namespace vtd {
namespace {

constexpr int MAXT = 64;

// returns (argmax, max probability) of one row; all lanes get the result
__device__ __forceinline__ void row_argmax(const float* __restrict__ row, int V, int is_prob, int lane, int* idx,
                                           float* pmax) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (is_prob) {
    for (int v = lane; v < V; v += 32) {
      float x = row[v];
      if (x > best) { best = x; bi = v; }
    }
  } else {
    // softmax exactly as written: exp(x - max) / sum, then argmax over the probabilities
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float ssum = 0.f;
    for (int v = lane; v < V; v += 32) ssum += expf(row[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    for (int v = lane; v < V; v += 32) {
      float p = __fdiv_rn(expf(row[v] - m), ssum);
      if (p > best) { best = p; bi = v; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  *idx = bi; *pmax = best;
}

__device__ __forceinline__ void collapse(const int* ids_t, const float* pm_t, int T, int V, int canonical,
                                         uint8_t* ids_out, int ids_cap, int* len_out, float* conf_out) {
  int prev = -1, len = 0;
  double csum = 0.0;
  (void)V;
  for (int t = 0; t < T; ++t) {
    int ci = ids_t[t];
    if (ci == 0) { if (canonical) prev = -1; continue; }
    if (ci == prev) continue;
    if (ci < 96) {                              // reverse_vocab.get(ci, '<unk>'): 96 and above are <unk>
      if (len < ids_cap) ids_out[len] = (uint8_t)ci;
      ++len;
      csum += (double)pm_t[len - 1];           // max(prediction[len(text)-1]); len-1 <= t always
    }
    prev = ci;
  }
  for (int k = len; k < ids_cap; ++k) ids_out[k] = 0;
  *len_out = len;
  *conf_out = len > 0 ? (float)(csum / (double)len) : 0.f;
}

__global__ void __launch_bounds__(1024) ctc_kernel(const float* __restrict__ x, int B, int T, int V, int ld, int is_prob,
                                                   int canonical, uint8_t* __restrict__ ids, int ids_stride,
                                                   int* __restrict__ lens, float* __restrict__ conf) {
  __shared__ int s_idx[MAXT];
  __shared__ float s_pm[MAXT];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int b = blockIdx.x;
  for (int t = wid; t < T; t += nw) {
    int bi; float pm;
    row_argmax(x + ((size_t)b * T + t) * ld, V, is_prob, lane, &bi, &pm);
    if (lane == 0) { s_idx[t] = bi; s_pm[t] = pm; }
  }
  __syncthreads();
  if (threadIdx.x == 0) collapse(s_idx, s_pm, T, V, canonical, ids + (size_t)b * ids_stride, ids_stride, lens + b, conf + b);
}

// same, but the result lands in the vtd_record of the crop (crop ci of the chunk <-> record via offsets)
__global__ void __launch_bounds__(1024) ctc_records_kernel(const float* __restrict__ x, int n_crops, int first_crop,
                                                           int T, int V, int ld, int canonical,
                                                           const int* __restrict__ offsets, int n, int kmax,
                                                           vtd_record* __restrict__ records) {
  __shared__ int s_idx[MAXT];
  __shared__ float s_pm[MAXT];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int b = blockIdx.x;
  if (first_crop + b >= offsets[n]) return;             // the launch may be sized for more crops than the batch holds
  for (int t = wid; t < T; t += nw) {
    int bi; float pm;
    row_argmax(x + ((size_t)b * T + t) * ld, V, 0, lane, &bi, &pm);
    if (lane == 0) { s_idx[t] = bi; s_pm[t] = pm; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ci = first_crop + b;
    int f = 0;
    while (f + 1 < n && offsets[f + 1] <= ci) ++f;
    vtd_record* r = records + (size_t)f * kmax + (ci - offsets[f]);
    int len; float cf;
    collapse(s_idx, s_pm, T, V, canonical, r->ids, 36, &len, &cf);
    r->len = len; r->rec_conf = cf;
  }
}

}  // namespace

cudaError_t ctc_greedy(const float* x, int B, int T, int V, int ld, int is_prob, int canonical, uint8_t* ids,
                       int ids_stride, int* lens, float* conf, cudaStream_t s, LaunchCounter* lc) {
  if (B <= 0) return cudaSuccess;
  if (T > MAXT || T <= 0 || V <= 1) return cudaErrorInvalidValue;
  ctc_kernel<<<B, 32 * (T < 32 ? T : 32), 0, s>>>(x, B, T, V, ld, is_prob, canonical, ids, ids_stride, lens, conf);
  if (lc) lc->n++;
  return cudaGetLastError();
}

cudaError_t ctc_into_records(const float* logits, int n_crops, int first_crop, int T, int V, int ld, int canonical,
                             const int* offsets, int n, int kmax, void* records, cudaStream_t s, LaunchCounter* lc) {
  if (n_crops <= 0) return cudaSuccess;
  if (T > MAXT || T <= 0) return cudaErrorInvalidValue;
  ctc_records_kernel<<<n_crops, 32 * (T < 32 ? T : 32), 0, s>>>(logits, n_crops, first_crop, T, V, ld, canonical, offsets, n, kmax,
                                                      reinterpret_cast<vtd_record*>(records));
  if (lc) lc->n++;
  return cudaGetLastError();
}

}  // namespace vtd
