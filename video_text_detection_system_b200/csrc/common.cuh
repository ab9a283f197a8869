// Shared declarations of libvtd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <string.h>

namespace vtd {

// The 16-bit storage type of the speed tier (activations, weights, tcgen05 A/B operands).  IEEE half in the shipped
// library: ten mantissa bits keep the probability / threshold maps within 2e-3 of the fp32 reference (north_star's bar
// for the 16-bit tier is 1e-2; bfloat16 measured 1.6e-2 / 2.0e-2 at 640x640, profiles/r01_bf16_error_budget.md), at the
// same tcgen05 kind::f16 rate and the same bytes.  -DVTD_BF16_STORAGE builds the same kernels over bfloat16 instead
// (libvtd_b200_bf16.so: fp32's range, for checkpoints whose activations exceed 65504).  The type keeps the name `bf16`
// throughout the sources in both variants.
#ifndef VTD_BF16_STORAGE
typedef __half bf16;
typedef __half2 bf16x2;
#define VTD_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define VTD_UMMA_AB_FMT 0u                              // instruction-descriptor A/B format: F16
__host__ __device__ __forceinline__ bf16 f32_to_16(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float f16_to_32(bf16 v) { return __half2float(v); }
__device__ __forceinline__ bf16x2 pack2(float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ float2 unpack2(bf16x2 h) { return __half22float2(h); }
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf16x2;
#define VTD_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define VTD_UMMA_AB_FMT 1u                              // instruction-descriptor A/B format: BF16
__host__ __device__ __forceinline__ bf16 f32_to_16(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float f16_to_32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16x2 pack2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 unpack2(bf16x2 h) { return __bfloat1622float2(h); }
#endif

// ---- element load/store helpers (activations are fp32 or bf16, math is fp32) -------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return f16_to_32(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return f32_to_16(v); }

struct LaunchCounter { long long n = 0; };

// Tuning / experiment switches (VTD_NO_*, VTD_TILE, VTD_DBG ...) exist only in -DVTD_DEV builds; the release library reads
// no environment variable on the compute path.
#ifdef VTD_DEV
inline const char* dev_env(const char* name) { return getenv(name); }
#define VTD_DBG_BITS(p) ((p).dbg)
#else
inline const char* dev_env(const char*) { return nullptr; }
#define VTD_DBG_BITS(p) 0
#endif

// Per-device one-time setup (cudaFuncSetAttribute is a per-device attribute; one process may hold contexts on several
// GPUs, and several host threads may reach a kernel's first launch at once): runs `fn` for the CURRENT device until it
// has succeeded once; the flag is published only AFTER fn returned, so a concurrent caller either sees it done or
// repeats the (idempotent) call itself -- it never launches ahead of the attribute.
struct PerDeviceFlag { unsigned long long mask[2] = {0ull, 0ull}; };
template <typename F>
inline cudaError_t once_per_device(PerDeviceFlag& f, F&& fn) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 127;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(&f.mask[dev >> 6], __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  const cudaError_t e = fn();
  if (e != cudaSuccess) return e;
  __atomic_fetch_or(&f.mask[dev >> 6], bit, __ATOMIC_RELEASE);
  return cudaSuccess;
}

// ---- implicit-GEMM convolution description (NHWC activations, [Cout][kh][kw][Cin] weights) ------
enum { RES_NONE = 0, RES_SAME = 1, RES_UP2 = 2 };          // residual add: same size / nearest-2x upsampled
enum { OUT_NHWC = 0, OUT_D2S = 1 };                        // plain store / depth-to-space 2x (ConvTranspose k2 s2)

struct ConvDesc {
  const void* in;      // [N,H,W,Cin]
  const void* w;       // [Cout][KH*KW*Cin]  (same element type as activations)
  const float* bias;   // [Cout] fp32 (folded BN)
  const void* res;     // residual, see res_mode (RES_SAME: [N,Ho,Wo,Cout]; RES_UP2: [N,Ho/2,Wo/2,Cout])
  void* out;           // OUT_NHWC: [N,Ho,Wo,Cout]; OUT_D2S: [N,2Ho,2Wo,Cout/4], channel co = (dy*2+dx)*(Cout/4)+c
  int N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
  int relu, res_mode, out_mode;   // relu: 0 none, 1 ReLU, 2 exact GELU (tcgen05 path only)
  int out_f32;         // 1: `out` is fp32 regardless of the activation type (logits, gate pre-activations)
  int hint_lw = -1, hint_lh = -1;   // tcgen05 path only: force the tile shape 2^lw x 2^lh pixels (GEMMs over token maps: whole rows)
  int pool = 0;        // tcgen05 path only: max-pool fused into the epilogue, `out` is the POOLED map: 1 = 2x2 s2, 2 = (2,1) s(2,1)
};

// generic CUDA-core path (any shape); T = float or bf16
template <typename T> cudaError_t conv_generic(const ConvDesc& d, cudaStream_t s, LaunchCounter* lc);

// tensor-core path (bf16, Cin%64==0, Cout%64==0, stride 1|2); returns cudaErrorNotSupported if shape unsupported
struct TcPlan;   // opaque: tensor maps + launch geometry, built once per layer
TcPlan* tc_plan_create(const ConvDesc& d, std::string* err);
TcPlan* tc_plan_create_win(const void* in, int N, int Hp, int Wp, int cpp, int stride, int nr, int Ho, int Wo,
                           const void* w, const float* bias, void* out, int relu, std::string* err, int pool = 0);
TcPlan* tc_plan_create_dbhead(const void* feat, int N, int H4, int W4, const void* w1, const float* b1_host,
                              const float* w2_host, const float* b2_host, float* prob, float* thresh, uint8_t* mask,
                              std::string* err);
TcPlan* tc_plan_create_lstm(const void* h_prev, void* h_next, int Bcap, const void* whh, const float* xproj, float* cbuf,
                            void* seq_out, int T, std::string* err);
cudaError_t dbhead_tcgen05(TcPlan* pl, int n, float thr, const float* logit_bias, cudaStream_t s, LaunchCounter* lc);
// one-pass DB head (3x3 convolutions of both branches + both transposed convolutions + sigmoid + mask in one kernel)
TcPlan* tc_plan_create_headfused(const ConvDesc& conv3x3_merged, const void* w1, const float* b1_host, const float* w2_host,
                                 const float* b2_host, float* prob, float* thresh, uint8_t* mask, std::string* err);
cudaError_t dbhead_fused_tcgen05(TcPlan* pl, int n, float thr, const float* logit_bias, cudaStream_t s, LaunchCounter* lc);
cudaError_t lstm_step_tcgen05(const TcPlan* pl, int B, int step, cudaStream_t s, LaunchCounter* lc);
void tc_plan_destroy(TcPlan*);

// persistent clustered BiLSTM layer (lstm_tcgen05.cu): one launch = all T steps of both directions
struct LstmPlan;
LstmPlan* lstm_plan_create(const void* whh /*[2*1024][256] bf16, rows (dir, unit tile, gate, unit)*/, std::string* err);
void lstm_plan_destroy(LstmPlan*);
cudaError_t bilstm_layer_tcgen05(const LstmPlan* pl, const void* xproj /*[B][T][2][1024] bf16*/, void* seq_out, int B, int T, cudaStream_t s,
                                 LaunchCounter* lc, const int* n_dyn = nullptr, int n_first = 0);
// DBNet stem (7x7 s2 + BN + ReLU) fused with the 3x3 s2 max-pool: the full-resolution stem map never reaches HBM
struct StemPoolPlan;
StemPoolPlan* stem_pool_plan_create(const void* in_padded, int N, int dh, int dw, const void* window_weights, const float* bias,
                                    void* pooled_out, std::string* err);
StemPoolPlan* stem_pool_plan_create_crnn(const void* crops_padded, int N, int crop_w, const void* window_weights, const float* bias,
                                         void* pooled_out, std::string* err);
void stem_pool_plan_destroy(StemPoolPlan*);
// n_dyn (optional, both): the launch is sized for n images but processes clamp(*n_dyn - n_first, 0, n) of them -- the crop count
// of a batch lives on the device, so the recogniser is launched without a host round trip
cudaError_t stem_pool_tcgen05(const StemPoolPlan* pl, int n, cudaStream_t s, LaunchCounter* lc, const int* n_dyn = nullptr,
                              int n_first = 0);
cudaError_t conv_tcgen05(const TcPlan* p, int n_actual, cudaStream_t s, LaunchCounter* lc, const int* n_dyn = nullptr, int n_first = 0);
bool tc_supported(const ConvDesc& d);

// ---- pooling -------------------------------------------------------------------------------------
template <typename T>
cudaError_t maxpool_nhwc(const T* in, T* out, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                         int ph, int pw, cudaStream_t s, LaunchCounter* lc);

// ---- preprocess ----------------------------------------------------------------------------------
struct ResizeTab {           // Pillow coefficient tables for one axis, device memory
  int in_size = 0, out_size = 0, ksize = 0;
  int maxcnt = 0;            // largest cnt[]: <= 4 selects the word-wise fast path of the preprocess kernel
  int* lo = nullptr;         // [out]
  int* cnt = nullptr;        // [out]
  int* kk = nullptr;         // [out][ksize]
};
// Placement of an NHWC image batch inside a (possibly zero-bordered) buffer, all in elements:
// element (n,y,x,c) lives at offset + n*img_pitch + y*row_pitch + x*cpp + c.
struct OutLayout { long long offset, img_pitch, row_pitch; int cpp; };
inline OutLayout dense_layout(int H, int W, int cpp) { return OutLayout{0, (long long)H * W * cpp, (long long)W * cpp, cpp}; }
inline OutLayout padded_layout(int H, int W, int cpp, int pt, int pb, int pl, int pr) {
  const long long row = (long long)(W + pl + pr) * cpp;
  return OutLayout{pt * row + (long long)pl * cpp, row * (H + pt + pb), row, cpp};
}
template <typename T>
cudaError_t preprocess_frames(const uint8_t* const* frames_dev /*device array of n pointers*/, int n, int h, int w,
                              int pitch, int pixfmt, const ResizeTab& tx, const ResizeTab& ty,
                              const float* lut /*[3][256], build_normalize_lut*/, T* out /*4 channels per pixel*/,
                              OutLayout lay, cudaStream_t s, LaunchCounter* lc);
cudaError_t build_normalize_lut(float* lut_dev /*768 floats*/, cudaStream_t s);

// ---- fused DB head tail ----------------------------------------------------------------------------
struct HeadTailWeights {     // both branches; fp32
  const float* w1;           // [2][256 (dy,dx,co)][64 ci]   folded BN
  const float* b1;           // [2][256]
  const float* w2;           // [2][64 co][4 (dy2,dx2)]
  const float* b2;           // [2]
};
template <typename T>
cudaError_t db_head_tail(const T* feat /*[N,H4,W4,128]*/, const HeadTailWeights& hw, int N, int H4, int W4,
                         const float* logit_bias /*[N,4H4,4W4] or null*/, float thr, float* prob, float* thresh,
                         uint8_t* mask, cudaStream_t s, LaunchCounter* lc);

// ---- box extraction ----------------------------------------------------------------------------------
struct BoxParams {
  int n, n_alloc, mh, mw;    // planes in this call / planes the workspace was laid out for / plane size
  int clip_h, clip_w;        // the reference's literal 640s
  int orig_h, orig_w;
  int kmax;
  float unclip;
};
struct BoxWorkLayout {       // byte offsets into one device workspace (see boxes.cu)
  size_t run_x, nruns, par, slot_of, comp, cand_slot, tmp, pool, zero_begin, comp_count, cand_count, pool_used,
      zero_end, overflow;
  int cap, kc, pool_words, cap_row;
};
size_t box_work_bytes(int n_alloc, int mh, int mw, int kc, BoxWorkLayout* lay);
cudaError_t extract_boxes(const float* prob, const uint8_t* mask, const BoxParams& p, uint8_t* work,
                          const BoxWorkLayout& lay, void* records /*vtd_record [n][kmax]*/, int* counts /*[n]*/,
                          cudaStream_t s, LaunchCounter* lc);
cudaError_t threshold_mask(const float* prob, uint8_t* mask, long long count, float thr, cudaStream_t s,
                           LaunchCounter* lc);

// ---- crop gather ---------------------------------------------------------------------------------------
cudaError_t scan_counts(const int* counts, int n, int* offsets /*[n+1]*/, cudaStream_t s, LaunchCounter* lc);
// crops [first_crop, first_crop+n_crops) of the batch (record order) -> out [n_crops,32,crop_w,4]
template <typename T>
cudaError_t crop_resize_records(const uint8_t* const* frames_dev, int src_h, int src_w, int pitch,
                                const void* records, const int* offsets, int n, int kmax, int first_crop,
                                int n_crops, int crop_w, int nv12 /*frames are NV12, converted on the fly*/, T* out,
                                OutLayout lay, cudaStream_t s, LaunchCounter* lc);
template <typename T>
cudaError_t crop_resize_list(const uint8_t* const* crops_dev, const int* h, const int* w, const int* pitch,
                             int n_crops, int crop_w, T* out, OutLayout lay, cudaStream_t s, LaunchCounter* lc);

// ---- LSTM ------------------------------------------------------------------------------------------------
// One layer, both directions. xproj: [B,T,2,4H] fp32 (= W_ih x + b_ih + b_hh, gate order i,f,g,o);
// whh: [2][4H][H]; out: [B,T,2H] (T type); hbuf: [2 ping-pong][2 dirs][B][H] fp32; cbuf: [2][B][H] fp32.
template <typename T, typename WT>
cudaError_t bilstm_layer(const float* xproj, const WT* whh, T* out, float* hbuf, float* cbuf, int B, int Tn,
                         int H, cudaStream_t s, LaunchCounter* lc);

// ---- CTC --------------------------------------------------------------------------------------------------
cudaError_t ctc_greedy(const float* x /*[B,T,ld], V <= ld*/, int B, int T, int V, int ld, int is_prob, int canonical,
                       uint8_t* ids /*[B][ids_stride]*/, int ids_stride, int* lens, float* conf, cudaStream_t s,
                       LaunchCounter* lc);
cudaError_t ctc_into_records(const float* logits, int n_crops, int first_crop, int T, int V, int ld, int canonical,
                             const int* offsets, int n, int kmax, void* records, cudaStream_t s, LaunchCounter* lc);

// ---- transformer recogniser (TrOCR branch, text_recognizer.py:39-69): the non-GEMM kernels (trocr.cu) -----------------------
cudaError_t layernorm_rows(const bf16* x, const float* gamma, const float* beta, bf16* y, long long rows, int C, float eps,
                           cudaStream_t s, LaunchCounter* lc);
cudaError_t attention_enc(const bf16* qkv /*[n][S][3D]*/, bf16* out /*[n][S][D]*/, int n, int S, int heads, float scale,
                          cudaStream_t s, LaunchCounter* lc);
cudaError_t attention_decode(const bf16* q, int ldq, const bf16* k, const bf16* v, int ldkv, int Lcap, int L, int n, int heads,
                             float scale, bf16* out /*[n][heads*64]*/, cudaStream_t s, LaunchCounter* lc, const int* tdev = nullptr,
                             int head_major = 0);
cudaError_t kv_to_head_major(const bf16* in /*[n][T][2][heads][64]*/, bf16* out /*[n][2][heads][T][64]*/, int n, int T, int heads,
                             cudaStream_t s, LaunchCounter* lc);
// tdev (optional, every decode-loop kernel below): the decode position lives in device memory, so that one captured CUDA
// graph of a decode step can be replayed for every position; argmax_rows advances it
cudaError_t vit_assemble(const bf16* patches, const bf16* cls, const bf16* pos, bf16* h, int n, int P, int D, cudaStream_t s,
                         LaunchCounter* lc);
cudaError_t trocr_embed(const int* ids, int ids_ld, int t, const bf16* tok, const bf16* pos, bf16* x, int n, int D, float scale,
                        cudaStream_t s, LaunchCounter* lc, const int* tdev = nullptr);
cudaError_t kv_append(const bf16* qkv /*[n][3D]*/, bf16* cache /*[n][Lcap][2D]*/, int n, int t, int Lcap, int D, cudaStream_t s,
                      LaunchCounter* lc, const int* tdev = nullptr);
cudaError_t argmax_rows(const float* logits, int n, int V, int ld, int* ids, int ids_ld, int t, int eos, int pad, int* finished,
                        int* n_finished, cudaStream_t s, LaunchCounter* lc, int* tdev = nullptr);
cudaError_t advance_position(int* tdev, cudaStream_t s, LaunchCounter* lc);      // *tdev += 1, after a step's argmax
// Linears of the decode loop (M <= 128 rows): every SM streams a slice of the weights (trocr.cu)
bool skinny_gemm_supported(int M, int N, int K);
cudaError_t skinny_gemm(const bf16* X, int ldx, const bf16* W, const float* bias, const bf16* res, int ldres, void* out, int ldo,
                        int out_f32, int M, int N, int K, int act, cudaStream_t s, LaunchCounter* lc);
struct TrocrCropMeta { int h, w, pitch, ksx, ksy, offx, offy; long long tmp_off; };   // == CropMeta of trocr.cu
cudaError_t trocr_resize_patches(const uint8_t* const* crops_dev, const void* meta_dev, const int* tab_dev, uint8_t* tmp, bf16* patches,
                                 int n, int S, int P, int max_h, cudaStream_t s, LaunchCounter* lc);
cudaError_t nchw_to_patches(const float* x, bf16* patches, int n, int S, int P, cudaStream_t s, LaunchCounter* lc);

// ---- annotated-frame overlay (overlay.cu) ----------------------------------------------------------------------
constexpr int OV_LABEL_MAX = 232;
struct OverlayItem { int32_t frame; int32_t bbox[4]; int32_t label_len; uint8_t label[OV_LABEL_MAX]; };   // == vtd_overlay_item
cudaError_t overlay_upload_tables(cudaStream_t s);
// items grouped by frame; frame_end[i] = one past the last item of item i's frame
cudaError_t draw_overlay(uint8_t* const* frames, int h, int w, int pitch, const OverlayItem* items, const int* frame_end,
                         int n_items, cudaStream_t s, LaunchCounter* lc);

// ---- layout helpers ----------------------------------------------------------------------------------------
template <typename T>
cudaError_t nchw_f32_to_nhwc(const float* in, T* out, int N, int C, int H, int W, OutLayout lay, cudaStream_t s,
                             LaunchCounter* lc);
template <typename T>
cudaError_t nhwc_to_nchw_f32(const T* in, float* out, int N, int C, int H, int W, OutLayout lay, cudaStream_t s,
                             LaunchCounter* lc);

}  // namespace vtd
