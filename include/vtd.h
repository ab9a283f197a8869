/*
 * vtd.h -- C ABI of libvtd_b200.so: the B200 (sm_100a) detect+recognize hot path of
 * malak29/video-text-detection-system.
 *
 * The reference has no FFI/plugin interface of its own (it is 100% Python); the boundary it
 * exposes to its Celery worker is the Python class surface of app/ml (TextDetector,
 * TextRecognizer, VideoTextPipeline, DBNet, CRNN).  The Python shims in
 * video_text_detection_system_b200/ keep that surface and bind the entry points below through
 * ctypes; each entry point names the reference code (file:line under /root/reference) whose
 * work it replaces.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: C linkage, opaque context, plain pointers and sizes, int status (0 = OK),
 * no torch types.  Every call is asynchronous on the context's CUDA stream unless it returns
 * data to a host pointer, in which case it synchronises that stream before returning.  A
 * context is bound to one device and is internally serialised by a mutex, so the reference's
 * 4 detect() threads (pipeliine.py:32,96-99) may share one.  There is no CPU fallback: every
 * entry point fails with VTD_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef VTD_H_
#define VTD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vtd_ctx vtd_ctx;

enum {
  VTD_OK = 0,
  VTD_ERR_ARG = 1,       /* bad argument / shape */
  VTD_ERR_CUDA = 2,      /* CUDA runtime/driver error (see vtd_last_error) */
  VTD_ERR_STATE = 3,     /* call out of order (e.g. detect before weights are loaded) */
  VTD_ERR_WEIGHT = 4,    /* state-dict entry missing or of the wrong shape */
  VTD_ERR_CAPACITY = 5   /* exceeds the sizes given at vtd_create */
};

enum { VTD_FP32 = 0,     /* fp32 activations, CUDA-core FFMA implicit GEMM: the <=1e-3 parity tier */
       VTD_16BIT = 1,    /* 16-bit activations, fp32 accumulate, tcgen05/TMEM implicit GEMM fed by TMA: the speed tier.  IEEE
                            half in libvtd_b200.so (maps within 2e-3 of the fp32 reference); libvtd_b200_bf16.so
                            (-DVTD_BF16_STORAGE, same sources and ABI) stores bfloat16 instead: fp32's range, ~8x the
                            rounding error (profiles/r01_bf16_error_budget.md) */
       VTD_BF16 = 1 };   /* round-1 name of VTD_16BIT */

enum { VTD_IDS_STRIDE = 64 }; /* row pitch of the ids_out arrays of the decode / recognise drop-ins */

enum { VTD_PIX_BGR = 0,  /* HxWx3 uint8, B,G,R interleaved (cv2 frames; text_detector.py:117-120) */
       VTD_PIX_NV12 = 1  /* H rows of Y then H/2 rows of interleaved UV (decoder surfaces) */ };

typedef struct vtd_config {
  int32_t device;         /* CUDA ordinal */
  int32_t backbone;       /* 18 or 50  (text_detector.py:16-20; resnet18 per BASELINE configs 1-4) */
  int32_t dtype;          /* VTD_FP32 | VTD_16BIT */
  int32_t det_h, det_w;   /* detector input size, multiples of 32; the reference hard-codes 640x640
                             (text_detector.py:101) */
  int32_t crop_w;         /* recogniser crop width: 128 = reference (text_recognizer.py:118), 100 = BASELINE cfg 3 */
  int32_t max_batch;      /* frames per call */
  int32_t max_boxes;      /* Kmax: boxes kept per frame (<= 1024) */
  int32_t max_src_h, max_src_w; /* largest source frame */
  int32_t canonical_ctc;  /* 0 = reference decode semantics (text_recognizer.py:151-163); 1 = canonical CTC */
  float   unclip_ratio;   /* 1.0 = off = reference behaviour; >1 grows each min-area rect by area*ratio/perimeter */
  int32_t flags;          /* VTD_FLAG_* (0 in production) */
  int32_t reserved[3];
} vtd_config;

/* Speed tier only: run the DB head as two kernels (3x3 convolutions -> feature map in HBM -> transposed convolutions)
 * instead of the one-pass kernel.  Same results bit for bit; keeps the intermediate map for vtd_debug_tensor("head"). */
enum { VTD_FLAG_UNFUSED_HEAD = 1,
/* Test aid: every device buffer of the context is allocated between two canary pages; vtd_check_guards() reports how
 * many canary bytes the kernels have overwritten (out-of-bounds WRITES next to a buffer; compute-sanitizer is not
 * available on every pool).  Costs 8 KB per buffer, nothing at run time. */
       VTD_FLAG_GUARD_ALLOCS = 2,
/* Speed tier only: run the DBNet stem and its 3x3 s2 max-pool as two kernels (the full-resolution stem map goes through
 * HBM) instead of the fused kernel.  Same results bit for bit. */
       VTD_FLAG_UNFUSED_STEM = 4 };

/* One entry of a PyTorch state dict, fp32, C-contiguous, host memory. */
typedef struct vtd_tensor {
  const char*  name;      /* reference key, e.g. "backbone.4.0.conv1.weight"  */
  const float* data;
  int32_t      ndim;
  int64_t      shape[4];
} vtd_tensor;

/* One detection, fixed 128 bytes (the unit gathered to rank 0 over NCCL). */
typedef struct vtd_record {
  int32_t frame;          /* index of the frame inside the batch */
  int32_t bbox[4];        /* x1,y1,x2,y2 in source-frame pixels   (text_detector.py:160-166) */
  int32_t polygon[8];     /* 4 x (x,y) in detector space, unscaled (text_detector.py:155,175) */
  float   det_conf;       /* mean probability inside the box       (text_detector.py:169-170) */
  float   rec_conf;       /* reference "confidence"                (text_recognizer.py:161-165) */
  int32_t len;            /* number of emitted token ids */
  uint8_t ids[36];        /* token ids 1..95 (text_recognizer.py:86-91); at most T<=31 used */
  int32_t start_index;    /* raster index of the component's first pixel (ordering key) */
  uint8_t pad[24];
} vtd_record;

/* ---- lifetime ------------------------------------------------------------------------------ */
int  vtd_create(vtd_ctx** out, const vtd_config* cfg);
void vtd_destroy(vtd_ctx* ctx);
const char* vtd_last_error(vtd_ctx* ctx);           /* ctx may be NULL: error of the last failed vtd_create */
int  vtd_set_stream(vtd_ctx* ctx, void* cuda_stream);  /* run on a caller-owned cudaStream_t (NULL = own stream) */
void* vtd_stream(vtd_ctx* ctx);
int  vtd_sync(vtd_ctx* ctx);
int64_t vtd_launch_count(vtd_ctx* ctx);             /* kernels launched by this context so far */
int  vtd_overflow_flag(vtd_ctx* ctx);               /* !=0: the last box extraction exceeded max_boxes / scratch */
int  vtd_time_T(vtd_ctx* ctx);                      /* CRNN sequence length T = crop_w/4 - 1 */
int  vtd_abi_version(void);
int  vtd_check_guards(vtd_ctx* ctx, int64_t* bad_bytes_out); /* VTD_FLAG_GUARD_ALLOCS contexts; synchronises the stream */

/* ---- weights: replaces load_state_dict (text_detector.py:106-113, text_recognizer.py:93-100).
 * Folds eval-mode BatchNorm (eps 1e-5) into the convolutions and repacks to the kernels' layouts. */
int  vtd_load_detector(vtd_ctx* ctx, const vtd_tensor* tensors, int n);
int  vtd_load_recognizer(vtd_ctx* ctx, const vtd_tensor* tensors, int n);

/* ---- stage 1: frame preprocessing.  Replaces cv2.cvtColor + ToPILImage/Resize/ToTensor/Normalize
 * (text_detector.py:99-104,117-124): BGR->RGB, Pillow antialiased bilinear resize (bit-exact,
 * 22-bit fixed point, u8 intermediate), /255, ImageNet mean/std, NHWC.  frames[i] points to frame i
 * (host or device memory); all n frames share h,w,pitch.  The source frames are kept (or copied to)
 * device memory for the later crop stage.  pixfmt = VTD_PIX_NV12: frames are decoder surfaces (h rows of Y, h/2 rows of
 * interleaved UV, even h and w, pitch in bytes); both this stage and the crop gather convert per tap with the BT.601
 * fixed point of cv2.cvtColor(COLOR_YUV2BGR_NV12), so the results equal those of the converted BGR frames bit for bit. */
int  vtd_preprocess(vtd_ctx* ctx, const uint8_t* const* frames, int n, int h, int w, int pitch,
                    int pixfmt, int frames_on_device);

/* ---- stages 2+3: DBNet ResNet+FPN conv stack and the fused DB head.  Replaces DBNet.forward
 * (text_detector.py:25-29 with the FPN repaired per SURVEY.md D5) + `> threshold` (:144).
 * Consumes the batch left by vtd_preprocess; leaves probability (fp32), threshold (fp32) and
 * mask (u8, prob > thr) planes, each [n, det_h, det_w], in device memory.  logit_bias_dev is an
 * optional device [n, det_h, det_w] fp32 plane added to the probability logit (NULL in production). */
int  vtd_detect_maps(vtd_ctx* ctx, int n, float thr, const float* logit_bias_dev);
int  vtd_get_maps(vtd_ctx* ctx, float** prob_dev, float** thresh_dev, uint8_t** mask_dev);
int  vtd_read_maps(vtd_ctx* ctx, int n, float* prob_host, float* thresh_host, uint8_t* mask_host); /* any may be NULL */

/* DBNet.forward drop-in for a caller-made tensor: x is [n,3,det_h,det_w] fp32 NCHW (host). */
int  vtd_dbnet_forward(vtd_ctx* ctx, const float* x_nchw_host, int n, float* prob_host, float* thresh_host);

/* ---- stage 4a: box extraction.  Replaces TextDetector._post_process (text_detector.py:143-178):
 * findContours(RETR_EXTERNAL) -> contourArea<100 reject -> minAreaRect -> boxPoints -> truncate ->
 * AABB clip -> scale -> size filter -> mean-probability confidence.  Works on the planes left by
 * vtd_detect_maps; records stay in device memory (vtd_read_records to fetch). */
int  vtd_extract_boxes(vtd_ctx* ctx, int n, int orig_h, int orig_w);

/* _post_process drop-in on a caller-supplied host map of any size mh x mw.  clip_h/clip_w are the
 * reference's literal 640s (:160-166,169-170).  out has room for cap records. */
int  vtd_postprocess_map(vtd_ctx* ctx, const float* prob_host, int mh, int mw, int clip_h, int clip_w,
                         int orig_w, int orig_h, float thr, vtd_record* out, int cap, int* n_out);

/* ---- stage 4b: crop gather + CRNN + CTC.  Replaces pipeliine.py:116-125 and
 * text_recognizer.py:114-167: axis-aligned crops of the ORIGINAL BGR frames, cv2.resize
 * INTER_LINEAR to 32 x crop_w, /255, CRNN conv stack, 2-layer BiLSTM, Linear, softmax, greedy decode
 * with the reference's collapse and confidence semantics. */
int  vtd_recognize_boxes(vtd_ctx* ctx, int n);

/* recognize_batch drop-in: n_crops host BGR crops (each h[i] x w[i] x 3, pitch[i] bytes per row).
 * ids_out is [n_crops][VTD_IDS_STRIDE]. logits_out may be NULL, else [n_crops, T, 97] fp32. */
int  vtd_recognize_crops(vtd_ctx* ctx, const uint8_t* const* crops, const int* h, const int* w,
                         const int* pitch, int n_crops, uint8_t* ids_out, int* len_out, float* conf_out,
                         float* logits_out);

/* CRNN.forward drop-in: x is [n,3,32,crop_w] fp32 NCHW host; logits [n,T,97] fp32 host. */
int  vtd_crnn_forward(vtd_ctx* ctx, const float* x_nchw_host, int n, float* logits_host);

/* _decode_prediction drop-in (text_recognizer.py:142-167): [B,T,V] fp32 host, probabilities
 * (is_prob=1) or logits (is_prob=0), T <= VTD_IDS_STRIDE.  ids_out is [B][VTD_IDS_STRIDE]. */
int  vtd_ctc_decode(vtd_ctx* ctx, const float* x_host, int B, int T, int V, int is_prob,
                    uint8_t* ids_out, int* len_out, float* conf_out);

/* ---- stage 4b', the reference's OTHER recogniser: TransformerRecognizer (text_recognizer.py:39-69), i.e. TrOCRProcessor +
 * VisionEncoderDecoderModel.generate(max_length=50), greedy.  Speed tier only.  Weights: the HuggingFace state dict of the
 * VisionEncoderDecoderModel (keys "encoder.*", "decoder.model.decoder.*", "decoder.output_projection.weight"; ViT encoder
 * with 64-wide heads, TrOCR decoder whose cross-attention reads the encoder width).  crops_per_chunk: crops processed
 * together (1..128; buffers are sized for it). */
int  vtd_load_trocr(vtd_ctx* ctx, const vtd_tensor* tensors, int n, int crops_per_chunk);
/* info8 = {image size, encoder tokens, encoder width, decoder width, vocabulary, max positions, crops per chunk, decoder layers} */
int  vtd_trocr_info(vtd_ctx* ctx, int32_t* info8);
/* recognize / recognize_batch drop-in: host BGR crops -> resize 384x384 (Pillow bilinear, as the processor), BGR->RGB, /255,
 * (x-0.5)/0.5 -> encoder -> greedy decode.  ids_out [n_crops][max_length] (int32, starts with the decoder start token, ends
 * with EOS, padded with the pad token as generate() pads); len_out [n_crops] (optional). */
int  vtd_trocr_generate_crops(vtd_ctx* ctx, const uint8_t* const* crops, const int* h, const int* w, const int* pitch,
                              int n_crops, int max_length, int32_t* ids_out, int* len_out);
/* Parity harness on processor output: pixel_values [n,3,S,S] fp32 host.  Any of the outputs may be NULL:
 * enc_out [n][tokens][encoder width] (last_hidden_state), logits_out [n][L][vocabulary] for the teacher-forced
 * decoder_ids [n][L], ids_out [n][max_length] / len_out [n] = greedy generate. */
int  vtd_trocr_forward(vtd_ctx* ctx, const float* pixel_values, int n, const int32_t* decoder_ids, int L, int max_length,
                       float* enc_out, float* logits_out, int32_t* ids_out, int* len_out);

/* ---- whole path in one call: preprocess -> detect -> boxes -> recognise (pipeliine.py:93-139).
 * records_host [n*max_boxes] / counts_host [n] may be NULL to leave results on the device. */
int  vtd_run_batch(vtd_ctx* ctx, const uint8_t* const* frames, int n, int h, int w, int pitch, int pixfmt,
                   int frames_on_device, float thr, const float* logit_bias_dev, int recognize,
                   vtd_record* records_host, int* counts_host);
int  vtd_read_records(vtd_ctx* ctx, int n, vtd_record* records_host, int* counts_host);
/* Device pointers of the results: records [max_batch*max_boxes], counts [max_batch].  The counts directly follow the
 * records (counts_dev == (int*)(records_dev + max_batch*max_boxes)), so a rank's results are one contiguous block of
 * max_batch*(max_boxes*128 + 4) bytes: one collective gathers them (parallel.gather_packed).
 * A batch with more than 1024 crops is recognised completely only once the host has looked at its crop count, which every
 * entry point that synchronises does (vtd_sync, vtd_read_records, the next vtd_run_batch): consume the device block after one
 * of them, or stream-ordered after vtd_run_batch when max_batch*max_boxes <= 1024. */
int  vtd_get_records(vtd_ctx* ctx, vtd_record** records_dev, int** counts_dev);

/* ---- result sink, annotated frames: ProcessingService._draw_detections drop-in (app/services/processing_service.py:188-218).
 * Draws, per item and in array order within a frame, the green 2-px bbox outline, the filled label plate above it and the
 * black label text (cv2.rectangle / cv2.getTextSize / cv2.putText, FONT_HERSHEY_SIMPLEX 0.5, thickness 1) into HxWx3 BGR
 * frames, IN PLACE: device frames are drawn where they are, host frames are copied up, drawn and copied back.  The label is
 * the caller's bytes ("%s (%.2f)" % (text, detection_confidence) in the reference, :198); bytes outside 32..126 draw '?'
 * as OpenCV does.  Items may be in any frame order; order among the items of one frame is the draw order.
 * frame index outside [0,n), label_len outside [0,VTD_OVERLAY_LABEL_MAX] or |coordinate| > 2^24: VTD_ERR_ARG. */
enum { VTD_OVERLAY_LABEL_MAX = 232 };
typedef struct vtd_overlay_item {
  int32_t frame;          /* index into frames[] */
  int32_t bbox[4];        /* x1,y1,x2,y2 as in the detection dict */
  int32_t label_len;
  uint8_t label[VTD_OVERLAY_LABEL_MAX];
} vtd_overlay_item;       /* 256 bytes */
int  vtd_draw_detections(vtd_ctx* ctx, uint8_t* const* frames, int n, int h, int w, int pitch, int frames_on_device,
                         const vtd_overlay_item* items, int n_items);

/* ---- parity harness: copy a named intermediate to the host as fp32 NCHW.
 * names: "input","c2","c3","c4","c5","p2_in","p2","head","crops","cnn","rnn0","rnn1","logits". */
int  vtd_debug_tensor(vtd_ctx* ctx, const char* name, int n, float* host_out, int64_t capacity, int64_t* shape4);

/* ---- measurement: per-launch device times of the conv / pool programs, taken with CUDA events on the
 * launching stream while the normal entry points run (bench.py's roofline figures).  which: 0 detector,
 * 1 recogniser.  info[16] = {kind (0 conv, 1 pool), tensor_core, H, W, Cin, Ho, Wo, Cout, KH, KW, stride,
 * launches timed, 0...}; *ms = summed device time of those launches.  which = 2: the other stages of the path, one
 * entry each: info[0] = 2, info[11] = launches timed, info[12] = stage (0 preprocess, 1 DB head tail, 2 box extraction,
 * 3 crop gather, 4/5 BiLSTM layer 0/1, 6 CTC decode). */
int  vtd_set_profiling(vtd_ctx* ctx, int on);
int  vtd_op_count(vtd_ctx* ctx, int which);
int  vtd_op_info(vtd_ctx* ctx, int which, int idx, int64_t* info, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* VTD_H_ */
