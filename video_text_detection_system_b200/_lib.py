"""ctypes binding of libvtd_b200.so (include/vtd.h) and a thin `Engine` wrapper.

There is no CPU fallback: if the shared object is missing it is built with nvcc when a compiler is
available, otherwise loading raises; if no sm_100 device is usable, creating an Engine raises VtdError.
"""
from __future__ import annotations

import ctypes as C
import gc
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvtd_b200.so")

VTD_FP32, VTD_16BIT = 0, 1
VTD_BF16 = VTD_16BIT                 # round-1 name of the enum value
VTD_PIX_BGR, VTD_PIX_NV12 = 0, 1
STAGE_NAMES = ("preprocess", "head_tail", "boxes", "crop", "lstm0", "lstm1", "ctc")   # vtd_op_info(which=2)
VTD_IDS_STRIDE = 64
VTD_FLAG_UNFUSED_HEAD = 1
VTD_FLAG_GUARD_ALLOCS = 2
VTD_FLAG_UNFUSED_STEM = 4

CHARS = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ "


class VtdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libvtd_b200 error %d: %s" % (code, msg))
        self.code = code


class VtdConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("backbone", C.c_int32), ("dtype", C.c_int32), ("det_h", C.c_int32),
                ("det_w", C.c_int32), ("crop_w", C.c_int32), ("max_batch", C.c_int32), ("max_boxes", C.c_int32),
                ("max_src_h", C.c_int32), ("max_src_w", C.c_int32), ("canonical_ctc", C.c_int32),
                ("unclip_ratio", C.c_float), ("flags", C.c_int32), ("reserved", C.c_int32 * 3)]


class VtdTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class VtdRecord(C.Structure):
    _fields_ = [("frame", C.c_int32), ("bbox", C.c_int32 * 4), ("polygon", C.c_int32 * 8), ("det_conf", C.c_float),
                ("rec_conf", C.c_float), ("len", C.c_int32), ("ids", C.c_uint8 * 36), ("start_index", C.c_int32),
                ("pad", C.c_uint8 * 24)]


RECORD_DTYPE = np.dtype([("frame", "<i4"), ("bbox", "<i4", (4,)), ("polygon", "<i4", (8,)), ("det_conf", "<f4"),
                         ("rec_conf", "<f4"), ("len", "<i4"), ("ids", "u1", (36,)), ("start_index", "<i4"),
                         ("pad", "u1", (24,))])
assert C.sizeof(VtdRecord) == 128 and RECORD_DTYPE.itemsize == 128

OVERLAY_LABEL_MAX = 232
OVERLAY_DTYPE = np.dtype([("frame", "<i4"), ("bbox", "<i4", (4,)), ("label_len", "<i4"), ("label", "u1", (OVERLAY_LABEL_MAX,))])
assert OVERLAY_DTYPE.itemsize == 256

_lib = None
_lib_lock = threading.Lock()

_u8pp = C.POINTER(C.c_void_p)


def exported_symbols() -> List[str]:
    """Every entry point include/vtd.h declares."""
    return ["vtd_create", "vtd_destroy", "vtd_last_error", "vtd_set_stream", "vtd_stream", "vtd_sync",
            "vtd_launch_count", "vtd_overflow_flag", "vtd_time_T", "vtd_abi_version", "vtd_check_guards", "vtd_load_detector",
            "vtd_load_recognizer", "vtd_preprocess", "vtd_detect_maps", "vtd_get_maps", "vtd_read_maps",
            "vtd_dbnet_forward", "vtd_extract_boxes", "vtd_postprocess_map", "vtd_recognize_boxes",
            "vtd_recognize_crops", "vtd_crnn_forward", "vtd_ctc_decode", "vtd_load_trocr", "vtd_trocr_info",
            "vtd_trocr_generate_crops", "vtd_trocr_forward", "vtd_run_batch", "vtd_read_records",
            "vtd_get_records", "vtd_draw_detections", "vtd_debug_tensor", "vtd_set_profiling", "vtd_op_count", "vtd_op_info"]


_libs: Dict[str, object] = {}


def load_library(variant: Optional[str] = None):
    """dlopen libvtd_b200.so (building it first if it is missing and nvcc exists).  variant: "" = the shipped library (IEEE
    half as the 16-bit storage type of the speed tier), "bf16" = libvtd_b200_bf16.so (bfloat16 storage, same ABI); None =
    what the environment variable VTD_STORAGE names, default "".  Both may be loaded in one process
    (Engine(dtype="fp16") beside Engine(dtype="bf16"))."""
    global _lib
    if variant is None:
        variant = os.environ.get("VTD_STORAGE", "")
    variant = "" if variant in ("f16", "fp16", "half") else variant
    with _lib_lock:
        if variant in _libs:
            return _libs[variant]
        from .build import build_library, variant_paths
        path = variant_paths(variant)[1]
        if not os.path.exists(path):
            build_library(variant=variant)
        lib = C.CDLL(path)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        lib.vtd_create.argtypes = [C.POINTER(vp), C.POINTER(VtdConfig)]
        lib.vtd_destroy.argtypes = [vp]; lib.vtd_destroy.restype = None
        lib.vtd_last_error.argtypes = [vp]; lib.vtd_last_error.restype = C.c_char_p
        lib.vtd_set_stream.argtypes = [vp, vp]
        lib.vtd_stream.argtypes = [vp]; lib.vtd_stream.restype = vp
        lib.vtd_sync.argtypes = [vp]
        lib.vtd_launch_count.argtypes = [vp]; lib.vtd_launch_count.restype = C.c_int64
        lib.vtd_overflow_flag.argtypes = [vp]
        lib.vtd_time_T.argtypes = [vp]
        lib.vtd_abi_version.argtypes = []
        lib.vtd_check_guards.argtypes = [vp, C.POINTER(C.c_int64)]
        lib.vtd_draw_detections.argtypes = [vp, _u8pp, i32, i32, i32, i32, i32, vp, i32]
        lib.vtd_load_detector.argtypes = [vp, C.POINTER(VtdTensor), i32]
        lib.vtd_load_recognizer.argtypes = [vp, C.POINTER(VtdTensor), i32]
        lib.vtd_preprocess.argtypes = [vp, _u8pp, i32, i32, i32, i32, i32, i32]
        lib.vtd_detect_maps.argtypes = [vp, i32, f32, vp]
        lib.vtd_get_maps.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        lib.vtd_read_maps.argtypes = [vp, i32, vp, vp, vp]
        lib.vtd_dbnet_forward.argtypes = [vp, vp, i32, vp, vp]
        lib.vtd_extract_boxes.argtypes = [vp, i32, i32, i32]
        lib.vtd_postprocess_map.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, i32, C.POINTER(i32)]
        lib.vtd_recognize_boxes.argtypes = [vp, i32]
        lib.vtd_recognize_crops.argtypes = [vp, _u8pp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp, vp, vp, vp]
        lib.vtd_crnn_forward.argtypes = [vp, vp, i32, vp]
        lib.vtd_ctc_decode.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp]
        lib.vtd_load_trocr.argtypes = [vp, C.POINTER(VtdTensor), i32, i32]
        lib.vtd_trocr_info.argtypes = [vp, C.POINTER(C.c_int32)]
        lib.vtd_trocr_generate_crops.argtypes = [vp, _u8pp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, i32, vp, vp]
        lib.vtd_trocr_forward.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]
        lib.vtd_run_batch.argtypes = [vp, _u8pp, i32, i32, i32, i32, i32, i32, f32, vp, i32, vp, vp]
        lib.vtd_read_records.argtypes = [vp, i32, vp, vp]
        lib.vtd_get_records.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        lib.vtd_debug_tensor.argtypes = [vp, C.c_char_p, i32, vp, C.c_int64, C.POINTER(C.c_int64)]
        lib.vtd_set_profiling.argtypes = [vp, i32]
        lib.vtd_op_count.argtypes = [vp, i32]
        lib.vtd_op_info.argtypes = [vp, i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        _libs[variant] = lib
        if _lib is None:
            _lib = lib
        return lib


def ids_to_text(ids: Sequence[int]) -> str:
    """text_recognizer.py:86-91 vocabulary: id i (1..95) -> CHARS[i-1]."""
    return "".join(CHARS[i - 1] for i in ids if 1 <= i <= len(CHARS))


def _state_dict_to_tensors(sd) -> Tuple[C.Array, list]:
    keep, items = [], []
    for k, v in sd.items():
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        if a.dtype.kind != "f" or a.ndim > 4:
            continue                      # num_batches_tracked etc.
        a = np.ascontiguousarray(a, dtype=np.float32)
        name = k.encode()
        keep.append((a, name))
        items.append((name, a))
    arr = (VtdTensor * len(items))()
    for i, (name, a) in enumerate(items):
        arr[i].name = name
        arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].ndim = a.ndim
        for d in range(a.ndim):
            arr[i].shape[d] = a.shape[d]
    return arr, keep


class Engine:
    """One vtd_ctx.  Thread-safe (the library serialises calls on a context)."""

    def __init__(self, device: int = 0, backbone: int = 18, dtype: str = "fp32", det_h: int = 640, det_w: int = 640,
                 crop_w: int = 128, max_batch: int = 1, max_boxes: int = 256, max_src_h: int = 2160,
                 max_src_w: int = 3840, canonical_ctc: bool = False, unclip_ratio: float = 1.0, fuse_head: bool = True, guard_allocs: bool = False, fuse_stem: bool = True):
        d = str(dtype).lower()
        # "fp16" / "bf16" name the 16-bit storage type of the speed tier and with it the library; "fp32" (the CUDA-core
        # parity tier) and "16bit" (the speed tier) take the process default (VTD_STORAGE, shipped = half)
        if d in ("fp16", "f16", "half", "float16"):
            variant = ""
        elif d in ("bf16", "bfloat16"):
            variant = "bf16"
        elif d in ("fp32", "float32", "0", "16bit", "1"):
            variant = None
        else:
            raise ValueError("dtype must be 'fp16', 'bf16', '16bit' or 'fp32', got %r" % (dtype,))
        self.lib = load_library(variant)
        speed = d not in ("fp32", "float32", "0")
        cfg = VtdConfig()
        cfg.device, cfg.backbone = int(device), int(backbone)
        cfg.dtype = VTD_16BIT if speed else VTD_FP32
        cfg.det_h, cfg.det_w, cfg.crop_w = int(det_h), int(det_w), int(crop_w)
        cfg.max_batch, cfg.max_boxes = int(max_batch), int(max_boxes)
        cfg.max_src_h, cfg.max_src_w = int(max_src_h), int(max_src_w)
        cfg.canonical_ctc = 1 if canonical_ctc else 0
        cfg.unclip_ratio = float(unclip_ratio)
        cfg.flags = 0 if fuse_head else VTD_FLAG_UNFUSED_HEAD      # parity harness: keep the "head" feature map (same results)
        if not fuse_stem:
            cfg.flags |= VTD_FLAG_UNFUSED_STEM                     # parity harness: stem and max-pool as two kernels (same results)
        if guard_allocs:
            cfg.flags |= VTD_FLAG_GUARD_ALLOCS                     # test aid: canary pages around every device buffer
        self.cfg = cfg
        self.det_h, self.det_w, self.crop_w = cfg.det_h, cfg.det_w, cfg.crop_w
        self.max_batch, self.max_boxes = cfg.max_batch, cfg.max_boxes
        stored = "bf16" if (variant == "bf16" or (variant is None and os.environ.get("VTD_STORAGE", "") == "bf16")) else "fp16"
        self.dtype = stored if speed else "fp32"
        self._h = C.c_void_p()
        rc = self.lib.vtd_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            raise VtdError(rc, (self.lib.vtd_last_error(None) or b"").decode())
        self.T = self.lib.vtd_time_T(self._h)
        self.det_loaded = self.rec_loaded = False

    # ---- plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise VtdError(rc, (self.lib.vtd_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.vtd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        self._check(self.lib.vtd_sync(self._h))

    def launch_count(self) -> int:
        return int(self.lib.vtd_launch_count(self._h))

    def stream(self) -> int:
        return int(self.lib.vtd_stream(self._h) or 0)

    def overflow(self) -> int:
        return int(self.lib.vtd_overflow_flag(self._h))

    def check_guards(self) -> int:
        """Canary bytes overwritten so far (guard_allocs=True contexts): 0 = no kernel wrote outside its buffers."""
        bad = C.c_int64(0)
        self._check(self.lib.vtd_check_guards(self._h, C.byref(bad)))
        return int(bad.value)

    # ---- weights
    def load_detector(self, state_dict):
        arr, keep = _state_dict_to_tensors(state_dict)
        self._check(self.lib.vtd_load_detector(self._h, arr, len(arr)))
        self.det_loaded = True

    def load_recognizer(self, state_dict):
        arr, keep = _state_dict_to_tensors(state_dict)
        self._check(self.lib.vtd_load_recognizer(self._h, arr, len(arr)))
        self.rec_loaded = True

    # ---- helpers
    @staticmethod
    def _frame_ptrs(frames, pixfmt: int = VTD_PIX_BGR) -> Tuple[C.Array, int, int, int, list]:
        """frames: sequence of HxWx3 uint8 arrays (same shape) or an [n,H,W,3] array; NV12: (H*3/2)xW uint8 planes."""
        keep = []
        ptrs = (C.c_void_p * len(frames))()
        h = w = pitch = None
        if pixfmt == VTD_PIX_NV12:
            for i, f in enumerate(frames):
                if not isinstance(f, np.ndarray) or f.dtype != np.uint8 or f.ndim != 2 or f.shape[0] % 3 or f.shape[1] % 2:
                    raise ValueError("NV12 frames must be (H*3/2)xW uint8 arrays with even H and W")
                f = np.ascontiguousarray(f)
                if h is None:
                    h, w, pitch = f.shape[0] * 2 // 3, f.shape[1], f.strides[0]
                elif (f.shape[0] * 2 // 3, f.shape[1]) != (h, w):
                    raise ValueError("all frames of a batch must have the same size")
                keep.append(f)
                ptrs[i] = f.ctypes.data
            return ptrs, h, w, pitch, keep
        for i, f in enumerate(frames):
            if not isinstance(f, np.ndarray) or f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise ValueError("frames must be HxWx3 uint8 arrays")
            if f.strides[2] != 1 or f.strides[1] != 3:
                f = np.ascontiguousarray(f)
            if h is None:
                h, w, pitch = f.shape[0], f.shape[1], f.strides[0]
            elif (f.shape[0], f.shape[1]) != (h, w) or f.strides[0] != pitch:
                if (f.shape[0], f.shape[1]) != (h, w):
                    raise ValueError("all frames of a batch must have the same size")
                f = np.ascontiguousarray(f)
                if f.strides[0] != pitch:
                    raise ValueError("all frames of a batch must share a row pitch")
            keep.append(f)
            ptrs[i] = f.ctypes.data
        return ptrs, h, w, pitch, keep

    # ---- stages
    def preprocess(self, frames, pixfmt: int = VTD_PIX_BGR):
        ptrs, h, w, pitch, keep = self._frame_ptrs(frames, pixfmt)
        self._check(self.lib.vtd_preprocess(self._h, C.cast(ptrs, _u8pp), len(frames), h, w, pitch, pixfmt, 0))
        self.sync()

    def preprocess_device(self, dev_ptrs: Sequence[int], h: int, w: int, pitch: int, pixfmt: int = VTD_PIX_BGR):
        ptrs = (C.c_void_p * len(dev_ptrs))(*[int(p) for p in dev_ptrs])
        self._check(self.lib.vtd_preprocess(self._h, C.cast(ptrs, _u8pp), len(dev_ptrs), h, w, pitch, pixfmt, 1))

    def detect_maps(self, n: int, thr: float = 0.5, logit_bias_dev: int = 0):
        self._check(self.lib.vtd_detect_maps(self._h, n, float(thr), C.c_void_p(logit_bias_dev or None)))

    def read_maps(self, n: int, prob=True, thresh=True, mask=True):
        shp = (n, self.det_h, self.det_w)
        p = np.empty(shp, np.float32) if prob else None
        t = np.empty(shp, np.float32) if thresh else None
        m = np.empty(shp, np.uint8) if mask else None
        self._check(self.lib.vtd_read_maps(self._h, n, p.ctypes.data if prob else None,
                                           t.ctypes.data if thresh else None, m.ctypes.data if mask else None))
        return p, t, m

    def dbnet_forward(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, self.det_h, self.det_w):
            raise ValueError("expected [n,3,%d,%d], got %s" % (self.det_h, self.det_w, x.shape))
        n = x.shape[0]
        p = np.empty((n, 1, self.det_h, self.det_w), np.float32)
        t = np.empty_like(p)
        self._check(self.lib.vtd_dbnet_forward(self._h, x.ctypes.data, n, p.ctypes.data, t.ctypes.data))
        return p, t

    def extract_boxes(self, n: int, orig_h: int, orig_w: int):
        self._check(self.lib.vtd_extract_boxes(self._h, n, int(orig_h), int(orig_w)))

    def recognize_boxes(self, n: int):
        self._check(self.lib.vtd_recognize_boxes(self._h, n))

    def read_records(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        rec = np.zeros((n, self.max_boxes), RECORD_DTYPE)
        cnt = np.zeros(n, np.int32)
        self._check(self.lib.vtd_read_records(self._h, n, rec.ctypes.data, cnt.ctypes.data))
        return rec, cnt

    def postprocess_map(self, prob: np.ndarray, orig_w: int, orig_h: int, thr: float, clip_h: Optional[int] = None,
                        clip_w: Optional[int] = None) -> np.ndarray:
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        if prob.ndim != 2:
            raise ValueError("probability map must be 2-D")
        rec = np.zeros(self.max_boxes, RECORD_DTYPE)
        n_out = C.c_int(0)
        self._check(self.lib.vtd_postprocess_map(self._h, prob.ctypes.data, prob.shape[0], prob.shape[1],
                                                 int(clip_h or self.det_h), int(clip_w or self.det_w), int(orig_w),
                                                 int(orig_h), float(thr), rec.ctypes.data, self.max_boxes,
                                                 C.byref(n_out)))
        return rec[:min(n_out.value, self.max_boxes)]

    def run_batch(self, frames, thr: float = 0.5, recognize: bool = True, logit_bias_dev: int = 0,
                  pixfmt: int = VTD_PIX_BGR, read: bool = True):
        ptrs, h, w, pitch, keep = self._frame_ptrs(frames, pixfmt)
        n = len(frames)
        rec = np.zeros((n, self.max_boxes), RECORD_DTYPE) if read else None
        cnt = np.zeros(n, np.int32) if read else None
        self._check(self.lib.vtd_run_batch(self._h, C.cast(ptrs, _u8pp), n, h, w, pitch, pixfmt, 0, float(thr),
                                           C.c_void_p(logit_bias_dev or None), 1 if recognize else 0,
                                           rec.ctypes.data if read else None, cnt.ctypes.data if read else None))
        return rec, cnt

    def run_batch_raw(self, ptrs, n: int, h: int, w: int, pitch: int, on_device: bool, thr: float, recognize: bool,
                      logit_bias_dev: int = 0, rec_ptr: int = 0, cnt_ptr: int = 0, pixfmt: int = VTD_PIX_BGR):
        """No allocation, no copies: caller owns every buffer (bench / multi-GPU driver)."""
        self._check(self.lib.vtd_run_batch(self._h, C.cast(ptrs, _u8pp), n, h, w, pitch, pixfmt, 1 if on_device else 0,
                                           float(thr), C.c_void_p(logit_bias_dev or None), 1 if recognize else 0,
                                           C.c_void_p(rec_ptr or None), C.c_void_p(cnt_ptr or None)))

    def device_records(self) -> Tuple[int, int]:
        r, c = C.c_void_p(), C.c_void_p()
        self._check(self.lib.vtd_get_records(self._h, C.byref(r), C.byref(c)))
        return int(r.value), int(c.value)

    def device_maps(self) -> Tuple[int, int, int]:
        p, t, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.vtd_get_maps(self._h, C.byref(p), C.byref(t), C.byref(m)))
        return int(p.value), int(t.value), int(m.value)

    def recognize_crops(self, crops: Sequence[np.ndarray], want_logits: bool = False):
        n = len(crops)
        keep = []
        ptrs = (C.c_void_p * n)()
        hs, ws, ps = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        for i, im in enumerate(crops):
            if not isinstance(im, np.ndarray) or im.ndim != 3 or im.shape[2] != 3 or im.size == 0:
                raise ValueError("crop %d is not a non-empty HxWx3 array" % i)
            if im.dtype != np.uint8:
                raise ValueError("crop %d is not uint8" % i)
            if im.strides[2] != 1 or im.strides[1] != 3 or im.strides[0] < 3 * im.shape[1]:
                im = np.ascontiguousarray(im)
            keep.append(im)
            ptrs[i], hs[i], ws[i], ps[i] = im.ctypes.data, im.shape[0], im.shape[1], im.strides[0]
        ids = np.zeros((n, VTD_IDS_STRIDE), np.uint8)
        lens = np.zeros(n, np.int32)
        conf = np.zeros(n, np.float32)
        logits = np.empty((n, self.T, 97), np.float32) if want_logits else None
        self._check(self.lib.vtd_recognize_crops(self._h, C.cast(ptrs, _u8pp), hs, ws, ps, n, ids.ctypes.data,
                                                 lens.ctypes.data, conf.ctypes.data,
                                                 logits.ctypes.data if want_logits else None))
        return ids, lens, conf, logits

    # ---- transformer recogniser (TrOCR branch)
    def load_trocr(self, state_dict, crops_per_chunk: int = 32):
        """state_dict: HuggingFace VisionEncoderDecoderModel.state_dict() (ViT encoder + TrOCR decoder)."""
        arr, keep = _state_dict_to_tensors(state_dict)
        self._check(self.lib.vtd_load_trocr(self._h, arr, len(arr), int(crops_per_chunk)))
        info = (C.c_int32 * 8)()
        self._check(self.lib.vtd_trocr_info(self._h, info))
        self.trocr = dict(zip(("image", "tokens", "enc_width", "dec_width", "vocab", "max_positions", "crops_per_chunk", "dec_layers"),
                              [int(v) for v in info]))

    def draw_detections(self, frames, items: np.ndarray, on_device: bool = False, h: int = 0, w: int = 0, pitch: int = 0) -> None:
        """vtd_draw_detections: draws `items` (OVERLAY_DTYPE records) into the frames IN PLACE.  frames: a list of HxWx3 uint8
        arrays of one size (host), or -- on_device -- a list of device addresses with h, w, pitch given."""
        items = np.ascontiguousarray(items, dtype=OVERLAY_DTYPE)
        n = len(frames)
        ptrs = (C.c_void_p * max(n, 1))()
        if on_device:
            for i, a in enumerate(frames):
                ptrs[i] = int(a)
        else:
            h, w = frames[0].shape[:2]
            pitch = frames[0].strides[0]
            for i, f in enumerate(frames):
                if (not isinstance(f, np.ndarray) or f.dtype != np.uint8 or f.ndim != 3 or f.shape != (h, w, 3) or f.strides[2] != 1
                        or f.strides[1] != 3 or f.strides[0] != pitch or not f.flags.writeable):
                    raise ValueError("frame %d: expected writeable HxWx3 uint8 frames of one size and pitch" % i)
                ptrs[i] = f.ctypes.data
        self._check(self.lib.vtd_draw_detections(self._h, C.cast(ptrs, _u8pp), n, int(h), int(w), int(pitch), 1 if on_device else 0,
                                                 items.ctypes.data, len(items)))

    def trocr_generate_crops(self, crops: Sequence[np.ndarray], max_length: int = 50) -> Tuple[np.ndarray, np.ndarray]:
        """BGR uint8 crops -> (ids [n, max_length] int32 padded as generate() pads, lengths [n])."""
        n = len(crops)
        keep = []
        ptrs = (C.c_void_p * n)()
        hs, ws, ps = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        for i, im in enumerate(crops):
            if not isinstance(im, np.ndarray) or im.ndim != 3 or im.shape[2] != 3 or im.size == 0 or im.dtype != np.uint8:
                raise ValueError("crop %d is not a non-empty HxWx3 uint8 array" % i)
            if im.strides[2] != 1 or im.strides[1] != 3 or im.strides[0] < 3 * im.shape[1]:
                im = np.ascontiguousarray(im)
            keep.append(im)
            ptrs[i], hs[i], ws[i], ps[i] = im.ctypes.data, im.shape[0], im.shape[1], im.strides[0]
        ids = np.zeros((n, max_length), np.int32)
        lens = np.zeros(n, np.int32)
        self._check(self.lib.vtd_trocr_generate_crops(self._h, C.cast(ptrs, _u8pp), hs, ws, ps, n, int(max_length),
                                                      ids.ctypes.data, lens.ctypes.data))
        return ids, lens

    def trocr_forward(self, pixel_values: np.ndarray, decoder_ids: Optional[np.ndarray] = None, max_length: int = 0,
                      want_encoder: bool = False):
        """Parity harness on processor output [n,3,S,S] fp32: returns (encoder states or None, teacher-forced logits
        [n,L,V] or None, greedy ids [n,max_length] or None, lengths or None)."""
        x = np.ascontiguousarray(pixel_values, dtype=np.float32)
        n = x.shape[0]
        info = self.trocr
        enc = np.empty((n, info["tokens"], info["enc_width"]), np.float32) if want_encoder else None
        logits = dids = None
        L = 0
        if decoder_ids is not None:
            dids = np.ascontiguousarray(decoder_ids, dtype=np.int32)
            L = dids.shape[1]
            logits = np.empty((n, L, info["vocab"]), np.float32)
        ids = np.zeros((n, max_length), np.int32) if max_length else None
        lens = np.zeros(n, np.int32) if max_length else None
        self._check(self.lib.vtd_trocr_forward(self._h, x.ctypes.data, n, dids.ctypes.data if dids is not None else None, L,
                                               int(max_length), enc.ctypes.data if enc is not None else None,
                                               logits.ctypes.data if logits is not None else None,
                                               ids.ctypes.data if ids is not None else None,
                                               lens.ctypes.data if lens is not None else None))
        return enc, logits, ids, lens

    def crnn_forward(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (3, 32, self.crop_w):
            raise ValueError("expected [n,3,32,%d], got %s" % (self.crop_w, x.shape))
        out = np.empty((x.shape[0], self.T, 97), np.float32)
        self._check(self.lib.vtd_crnn_forward(self._h, x.ctypes.data, x.shape[0], out.ctypes.data))
        return out

    def ctc_decode(self, x: np.ndarray, is_prob: bool):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 2:
            x = x[None]
        B, T, V = x.shape
        ids = np.zeros((B, VTD_IDS_STRIDE), np.uint8)
        lens = np.zeros(B, np.int32)
        conf = np.zeros(B, np.float32)
        self._check(self.lib.vtd_ctc_decode(self._h, x.ctypes.data, B, T, V, 1 if is_prob else 0, ids.ctypes.data,
                                            lens.ctypes.data, conf.ctypes.data))
        return ids, lens, conf

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.vtd_set_stream(self._h, C.c_void_p(cuda_stream or None)))

    def set_profiling(self, on: bool):
        self._check(self.lib.vtd_set_profiling(self._h, 1 if on else 0))

    def op_profile(self, which: int) -> List[Dict]:
        """Per-op shapes and summed device time (ms) of the launches timed since set_profiling(True)."""
        out = []
        for i in range(self.lib.vtd_op_count(self._h, which)):
            info = (C.c_int64 * 16)()
            ms = C.c_double(0.0)
            self._check(self.lib.vtd_op_info(self._h, which, i, info, C.byref(ms)))
            keys = ("kind", "tensor_core", "H", "W", "Cin", "Ho", "Wo", "Cout", "KH", "KW", "stride", "launches")
            d = {k: int(info[j]) for j, k in enumerate(keys)}
            d["ms"] = float(ms.value)
            d["index"] = i
            if which == 2:
                d["name"] = STAGE_NAMES[int(info[12])]
            out.append(d)
        return out

    def debug_tensor(self, name: str, n: int) -> np.ndarray:
        shape = (C.c_int64 * 4)()
        self._check(self.lib.vtd_debug_tensor(self._h, name.encode(), n, None, 0, shape))
        out = np.empty(tuple(int(s) for s in shape), np.float32)
        self._check(self.lib.vtd_debug_tensor(self._h, name.encode(), n, out.ctypes.data, out.size, shape))
        return out


# id -> character as a bytes.translate table: ids 1..95 map to CHARS (printable ASCII), every other byte to NUL
_ID_TABLE = bytes(ord(CHARS[i - 1]) if 1 <= i <= len(CHARS) else 0 for i in range(256))
_IDS_BYTES = 36                      # sizeof(vtd_record.ids)


class _GcPaused:
    """Pause the cyclic collector while a batch of result dictionaries is built.  A 16-frame batch is ~5 600 new
    containers (dict + bbox + polygon lists per detection), none of them cyclic; with the collector running, their
    allocation triggers young-generation passes and costs several times the conversion itself (6.9 ms vs 1.3 ms per
    batch, measured).  Re-entrant and thread-safe: the collector is switched back on when the last concurrent
    user leaves, and only if it was on when the first one entered."""
    _lock, _depth, _was_enabled = threading.Lock(), 0, False

    def __enter__(self):
        cls = _GcPaused
        with cls._lock:
            if cls._depth == 0:
                cls._was_enabled = gc.isenabled()
                gc.disable()
            cls._depth += 1

    def __exit__(self, *exc):
        cls = _GcPaused
        with cls._lock:
            cls._depth -= 1
            if cls._depth == 0 and cls._was_enabled:
                gc.enable()
        return False


gc_paused = _GcPaused


def _texts(r: np.ndarray, count: int) -> List[str]:
    """Texts of `count` records in one pass: the whole ids column is mapped with one bytes.translate and decoded
    once (C speed, no per-character Python work); a record's text is then a slice of `len` characters.  Ids outside
    1..95 (blank, <unk>, padding) become NULs and are dropped, as ids_to_text drops them."""
    lens = np.minimum(r["len"], _IDS_BYTES).tolist()
    txt = np.ascontiguousarray(r["ids"]).tobytes().translate(_ID_TABLE).decode("ascii")
    out = [txt[o:o + n] for o, n in zip(range(0, count * _IDS_BYTES, _IDS_BYTES), lens)]
    if "\0" in txt:                  # only padding beyond `len` in the normal case: the slices hold none of it
        out = [t.replace("\0", "") if "\0" in t else t for t in out]
    return out


def records_to_detections(rec_row: np.ndarray, count: int, with_text: bool) -> List[Dict]:
    """vtd_record rows of one frame -> the reference's detection dicts (plain Python scalars).  Field extraction is
    vectorised (one .tolist() per field), so assembling ~50 dicts per frame costs microseconds, not milliseconds."""
    r = rec_row[:count]
    if count <= 0:
        return []
    bbox = r["bbox"].tolist()
    conf = r["det_conf"].astype(np.float64).tolist()
    poly = r["polygon"].reshape(count, 4, 2).tolist()
    if not with_text:
        return [{"bbox": b, "confidence": c, "polygon": p} for b, c, p in zip(bbox, conf, poly)]
    lens = np.minimum(r["len"], _IDS_BYTES).tolist()
    ids = r["ids"].tolist()
    rconf = r["rec_conf"].astype(np.float64).tolist()
    return [{"bbox": b, "confidence": c, "polygon": p, "ids": i[:n], "text": t, "recognition_confidence": rc}
            for b, c, p, i, n, t, rc in zip(bbox, conf, poly, ids, lens, _texts(r, count), rconf)]


def records_to_regions(rec_row: np.ndarray, count: int) -> List[Dict]:
    """vtd_record rows of one frame -> the pipeline's text regions (pipeliine.py:127-133), built in one pass."""
    r = rec_row[:count]
    if count <= 0:
        return []
    return [{"bbox": b, "text": t, "detection_confidence": c, "recognition_confidence": rc, "polygon": p}
            for b, t, c, rc, p in zip(r["bbox"].tolist(), _texts(r, count), r["det_conf"].astype(np.float64).tolist(),
                                      r["rec_conf"].astype(np.float64).tolist(),
                                      r["polygon"].reshape(count, 4, 2).tolist())]
