"""CPU oracle for the detect+recognize hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch fp32 eager + Pillow + OpenCV + NumPy,
i.e. the very third-party arithmetic the reference calls) of the reference's
per-frame path.  It exists so the CUDA path can be checked on a machine where
/root/reference is not mounted (the GPU box).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; nothing in
video_text_detection_system_b200/ does.

Parity status: the reference's own tests hold NO golden vectors or numeric
assertions for this path, so the oracle is pinned the
other way: oracle/check_against_reference.py imports the reference's source
files in the build container and compares every function here against them on
seeded inputs (exact equality), and oracle/make_goldens.py writes the
reference's own outputs to tests/golden/*.npz.

Reference (paths relative to /root/reference):
  app/ml/models/text_detector.py    DBNet :12-29, FeaturePyramidNetwork :31-56,
                                    DBHead :58-86, TextDetector :88-178
  app/ml/models/text_recognizer.py  CRNN :12-37, TextRecognizer :71-167
  app/ml/inference/pipeliine.py     crop + recognise loops :104-139, :143-172

Deviations from the reference as shipped (all forced):
  * FPN wiring: lateral i is fed backbone level C(5-i) (the shipped forward
    feeds C5 to every lateral and cannot run).  State-dict keys are unchanged.
  * backbone may be resnet18 (in_channels=512) as BASELINE configs 1-4 ask.
  * pretrained=False (no network).
  * np.int0 -> astype(np.intp) (removed in NumPy 2; same truncation).
  * the literal 640 of text_detector.py:101,160-170 is generalised to
    (det_h, det_w); at 640x640 every formula reduces to the reference verbatim.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch
import torch.nn as nn
import torchvision

from video_text_detection_system_b200.synthetic import planted_logit_bias, randomize_bn, synthetic_frames  # noqa: F401

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
CHARS = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ "


# ----------------------------------------------------------------------------
# Detector network (text_detector.py:12-86)
# ----------------------------------------------------------------------------
class _FPN(nn.Module):
    """text_detector.py:31-40 (constructor identical; forward lives in dbnet_forward)."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.inner_blocks = nn.ModuleList(
            [nn.Conv2d(in_channels // (2 ** i), 256, 1) for i in range(4)])
        self.layer_blocks = nn.ModuleList(
            [nn.Conv2d(256, 256, 3, padding=1) for _ in range(4)])


def _head_branch(c: int) -> nn.Sequential:
    """One branch of DBHead, text_detector.py:61-70."""
    q = c // 4
    return nn.Sequential(
        nn.Conv2d(c, q, 3, padding=1), nn.BatchNorm2d(q), nn.ReLU(inplace=True),
        nn.ConvTranspose2d(q, q, 2, stride=2), nn.BatchNorm2d(q), nn.ReLU(inplace=True),
        nn.ConvTranspose2d(q, 1, 2, stride=2), nn.Sigmoid())


class _Head(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.probability_head = _head_branch(c)
        self.threshold_head = _head_branch(c)


class OracleDBNet(nn.Module):
    """Same parameter names/shapes as the reference DBNet (Appendix D of SURVEY.md)."""

    def __init__(self, backbone: str = "resnet18"):
        super().__init__()
        if backbone == "resnet50":
            rn, cin = torchvision.models.resnet50(weights=None), 2048
        elif backbone == "resnet18":
            rn, cin = torchvision.models.resnet18(weights=None), 512
        else:
            raise ValueError(backbone)
        self.backbone = nn.Sequential(*list(rn.children())[:-2])   # text_detector.py:19
        self.fpn = _FPN(cin)                                       # :22
        self.head = _Head(256)                                     # :23

    def forward(self, x, logit_bias=None):
        return dbnet_forward(self, x, logit_bias=logit_bias)


def dbnet_forward(net, x: torch.Tensor, logit_bias: Optional[torch.Tensor] = None,
                  return_feats: bool = False):
    """DBNet.forward (text_detector.py:25-29) with the repaired FPN (:42-56).

    `net` may be an OracleDBNet or the reference's own DBNet instance.
    logit_bias: optional [N,1,H,W] plane added to the probability head's
    pre-sigmoid logit (benchmark config 3's planted map); None in
    production and in every reference-parity case.
    """
    b = net.backbone
    x = b[3](b[2](b[1](b[0](x))))
    c2 = b[4](x)
    c3 = b[5](c2)
    c4 = b[6](c3)
    c5 = b[7](c4)
    feats = (c5, c4, c3, c2)
    fpn = net.fpn
    last = fpn.inner_blocks[0](feats[0])
    for i in range(1, 4):
        lateral = fpn.inner_blocks[i](feats[i])
        top = nn.functional.interpolate(last, scale_factor=2, mode="nearest")
        last = lateral + top
    p2 = fpn.layer_blocks[3](last)              # only results[-1] is returned (:56)
    ph, th = net.head.probability_head, net.head.threshold_head
    if logit_bias is None:
        prob = ph(p2)
    else:
        prob = torch.sigmoid(ph[:-1](p2) + logit_bias)
    thr = th(p2)
    out = {"probability": prob, "threshold": thr}
    if return_feats:
        out.update(c2=c2, c3=c3, c4=c4, c5=c5, p2_in=last, p2=p2)
    return out


def build_dbnet(backbone: str = "resnet18", seed: int = 0, random_bn: bool = True) -> OracleDBNet:
    torch.manual_seed(seed)
    net = OracleDBNet(backbone)
    if random_bn:
        randomize_bn(net, seed + 1000)
    return net.eval()


# ----------------------------------------------------------------------------
# Detector preprocessing (text_detector.py:99-104, :117-124)
# ----------------------------------------------------------------------------
def make_transform(det_h: int = 640, det_w: int = 640):
    import torchvision.transforms as T
    return T.Compose([T.ToPILImage(), T.Resize((det_h, det_w)), T.ToTensor(),
                      T.Normalize(mean=list(IMAGENET_MEAN), std=list(IMAGENET_STD))])


def preprocess(image: np.ndarray, det_h: int = 640, det_w: int = 640) -> torch.Tensor:
    """text_detector.py:119-124 -> [1,3,det_h,det_w] fp32 (RGB, ImageNet-normalised)."""
    rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB) if image.ndim == 3 else image
    return make_transform(det_h, det_w)(rgb).unsqueeze(0)


def pillow_coeffs(in_size: int, out_size: int):
    """Pillow ImagingResample precompute_coeffs + normalize_coeffs_8bpc for BILINEAR
    (support 1.0, antialias when down-scaling).  Returns (xmin[out], count[out],
    kk[out, ksize] int32) with 22-bit fixed-point weights."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        for x in range(n):
            a = abs((x + lo - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        q = np.where(w < 0, (-0.5 + w * (1 << 22)).astype(np.int64), (0.5 + w * (1 << 22)).astype(np.int64))
        xmin[xx], cnt[xx] = lo, n
        kk[xx, :n] = q
    return xmin, cnt, kk


def pillow_resize_restated(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """NumPy restatement of PIL.Image.resize((out_w,out_h), BILINEAR) on HxWxC uint8:
    horizontal pass first, uint8 intermediate, then vertical.
    This is the arithmetic the CUDA preprocess kernel implements.

    Domain: bit-exact to Pillow 12.2 for every source no taller than 100x its width
    (tests/test_oracle_golden.py sweeps sizes, up- and down-scaling on either axis).  For a source with
    h > 100*w whose height shrinks, Pillow runs the vertical pass first and the u8 intermediate rounds
    differently (+-1 LSB); video frames are nowhere near that aspect ratio."""
    h, w, c = img.shape
    src = img.astype(np.int64)
    if out_w != w:
        xmin, cnt, kk = pillow_coeffs(w, out_w)
        tmp = np.empty((h, out_w, c), np.int64)
        for xx in range(out_w):
            n = cnt[xx]
            acc = (src[:, xmin[xx]:xmin[xx] + n, :] * kk[xx, :n][None, :, None].astype(np.int64)).sum(1)
            tmp[:, xx, :] = np.clip((acc + (1 << 21)) >> 22, 0, 255)
        src = tmp
    if out_h != h:
        ymin, cnt, kk = pillow_coeffs(h, out_h)
        out = np.empty((out_h, src.shape[1], c), np.int64)
        for yy in range(out_h):
            n = cnt[yy]
            acc = (src[ymin[yy]:ymin[yy] + n] * kk[yy, :n][:, None, None].astype(np.int64)).sum(0)
            out[yy] = np.clip((acc + (1 << 21)) >> 22, 0, 255)
        src = out
    return src.astype(np.uint8)


def preprocess_restated(image: np.ndarray, det_h: int, det_w: int) -> np.ndarray:
    """Same result as preprocess() but through the restated integer resize;
    returns [3,det_h,det_w] fp32."""
    rgb = image[:, :, ::-1]
    r = pillow_resize_restated(np.ascontiguousarray(rgb), det_h, det_w)
    x = r.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(IMAGENET_MEAN, np.float32)) / np.asarray(IMAGENET_STD, np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


# ----------------------------------------------------------------------------
# Detector post-processing (text_detector.py:143-178)
# ----------------------------------------------------------------------------
def unclip_rect(rect, ratio: float):
    """north_star's "unclip" (an extension: the reference has none): grow a min-area rect by the
    DB offset d = area * ratio / perimeter on every side, float32 op for op as csrc/box_geom.cuh unclip_rect."""
    (cx, cy), (w, h), ang = rect
    w, h, r = np.float32(w), np.float32(h), np.float32(ratio)
    per = np.float32(2.0) * (w + h)
    if not (ratio > 1.0) or not (per > 0):
        return rect
    d = (w * h) * r / per
    two_d = np.float32(2.0) * d
    return (cx, cy), (float(w + two_d), float(h + two_d)), ang


def post_process(prob_map: np.ndarray, orig_width: int, orig_height: int, threshold: float,
                 det_h: Optional[int] = None, det_w: Optional[int] = None, unclip_ratio: float = 1.0) -> List[Dict]:
    """_post_process with 640 -> (det_h, det_w).  The reference hard-codes 640 even
    when the map has another size (tests/test_models.py:51 passes 160x160); that
    behaviour is reproduced when det_h/det_w are left None (=640).  unclip_ratio = 1.0
    (the default) is the reference verbatim."""
    det_h = 640 if det_h is None else det_h
    det_w = 640 if det_w is None else det_w
    binary_map = (prob_map > threshold).astype(np.uint8) * 255            # :144 strict >
    contours, _ = cv2.findContours(binary_map, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    detections = []
    for contour in contours:
        if cv2.contourArea(contour) < 100:                                  # :150
            continue
        rect = cv2.minAreaRect(contour)
        if unclip_ratio > 1.0:
            rect = unclip_rect(rect, unclip_ratio)
        box = cv2.boxPoints(rect).astype(np.intp)                           # :153-155 (np.int0)
        xs, ys = box[:, 0], box[:, 1]
        x1, y1 = max(0, int(np.min(xs))), max(0, int(np.min(ys)))           # :160
        x2, y2 = min(det_w, int(np.max(xs))), min(det_h, int(np.max(ys)))   # :161
        x1 = int(x1 * orig_width / det_w)                                   # :163-166
        y1 = int(y1 * orig_height / det_h)
        x2 = int(x2 * orig_width / det_w)
        y2 = int(y2 * orig_height / det_h)
        if x2 - x1 > 10 and y2 - y1 > 10:                                   # :168
            confidence = float(np.mean(prob_map[y1 * det_h // orig_height:y2 * det_h // orig_height,
                                                x1 * det_w // orig_width:x2 * det_w // orig_width]))
            detections.append({"bbox": [x1, y1, x2, y2], "confidence": confidence,
                               "polygon": box.tolist()})
    return detections


def detect(net: nn.Module, image: np.ndarray, confidence_threshold: float = 0.5,
           det_h: int = 640, det_w: int = 640, logit_bias: Optional[torch.Tensor] = None) -> List[Dict]:
    """TextDetector.detect, text_detector.py:115-141 (never raises)."""
    try:
        oh, ow = image.shape[:2]
        x = preprocess(image, det_h, det_w)
        with torch.no_grad():
            prob = dbnet_forward(net, x, logit_bias)["probability"].cpu().numpy()[0, 0]
        return post_process(prob, ow, oh, confidence_threshold, det_h, det_w)
    except Exception:
        return []


# ----------------------------------------------------------------------------
# Recognizer (text_recognizer.py:12-37, :86-167)
# ----------------------------------------------------------------------------
class OracleCRNN(nn.Module):
    """text_recognizer.py:12-37; same parameter names."""

    def __init__(self, vocab_size: int, hidden_size: int = 256, num_layers: int = 2):
        super().__init__()
        L = []
        cfg = [(3, 64, 3, 1, (2, 2)), (64, 128, 3, 1, (2, 2)), (128, 256, 3, 1, None),
               (256, 256, 3, 1, ((2, 1), (2, 1))), (256, 512, 3, 1, None),
               (512, 512, 3, 1, ((2, 1), (2, 1))), (512, 512, 2, 0, None)]
        for cin, cout, k, p, pool in cfg:
            L += [nn.Conv2d(cin, cout, k, 1, p), nn.BatchNorm2d(cout), nn.ReLU(True)]
            if pool is not None:
                L.append(nn.MaxPool2d(*pool))
        self.cnn = nn.Sequential(*L)
        self.rnn = nn.LSTM(512, hidden_size, num_layers, batch_first=True, bidirectional=True)
        self.classifier = nn.Linear(hidden_size * 2, vocab_size)

    def forward(self, x):
        f = self.cnn(x)
        b, c, h, w = f.size()
        f = f.view(b, c * h, w).permute(0, 2, 1)          # :32
        r, _ = self.rnn(f)
        return self.classifier(r)


def build_vocab() -> Dict[str, int]:
    """text_recognizer.py:86-91."""
    vocab = {ch: i + 1 for i, ch in enumerate(CHARS)}
    vocab["<blank>"] = 0
    vocab["<unk>"] = len(vocab)
    return vocab


def build_crnn(seed: int = 0, random_bn: bool = True) -> OracleCRNN:
    torch.manual_seed(seed)
    net = OracleCRNN(len(build_vocab()))
    if random_bn:
        randomize_bn(net, seed + 2000)
    return net.eval()


def crnn_inputs(images: Sequence[np.ndarray], crop_w: int = 128) -> torch.Tensor:
    """text_recognizer.py:116-122: cv2.resize(img,(128,32)) INTER_LINEAR, HWC->CHW, /255,
    BGR order kept."""
    ts = []
    for img in images:
        r = cv2.resize(img, (crop_w, 32))
        ts.append(torch.from_numpy(r).permute(2, 0, 1).float() / 255.0)
    return torch.stack(ts)


def cv_resize_linear_restated(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Restatement of OpenCV's portable C++ INTER_LINEAR u8 path:
    half-pixel centres, 11-bit coefficients, edge clamp.  The installed wheel may
    differ from this by 1 LSB on a small share of pixels (IPP path)."""
    h, w = img.shape[:2]
    c = img.shape[2] if img.ndim == 3 else 1
    src = img.reshape(h, w, c).astype(np.int64)

    def tab(n_in, n_out):
        idx = np.zeros(n_out, np.int64)
        a = np.zeros((n_out, 2), np.int64)
        sc = n_in / n_out
        for d in range(n_out):
            f = np.float32((d + 0.5) * sc - 0.5)
            s = int(math.floor(f))
            f = np.float32(f - s)
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= n_in - 1:
                s, f = n_in - 1, np.float32(0)
            idx[d] = s
            a1 = int(np.rint(np.float32(f) * np.float32(2048)))
            a0 = int(np.rint((np.float32(1.0) - np.float32(f)) * np.float32(2048)))
            a[d] = (a0, a1)
        return idx, a

    xi, xa = tab(w, out_w)
    yi, ya = tab(h, out_h)
    xi1 = np.minimum(xi + 1, w - 1)
    rows = src[:, xi, :] * xa[:, 0][None, :, None] + src[:, xi1, :] * xa[:, 1][None, :, None]
    yi1 = np.minimum(yi + 1, h - 1)
    r0, r1 = rows[yi], rows[yi1]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out.reshape(out_h, out_w, c) if img.ndim == 3 else out.reshape(out_h, out_w)


def decode_prediction(prediction: torch.Tensor, vocab: Optional[Dict[str, int]] = None
                      ) -> Tuple[str, float, List[int]]:
    """_decode_prediction, text_recognizer.py:142-167, plus the emitted token ids.

    Reference semantics (not canonical CTC): blanks do not reset prev_char; <unk> is
    dropped from the text but still becomes prev_char; the confidence of the k-th
    emitted character is max(prediction[k-1]) -- indexed by emitted count, not time."""
    vocab = vocab or build_vocab()
    pred_indices = torch.argmax(prediction, dim=1)
    reverse_vocab = {v: k for k, v in vocab.items()}
    text, ids, confidences, prev = "", [], [], None
    for idx in pred_indices:
        ci = idx.item()
        if ci == 0:
            continue
        if ci == prev:
            continue
        ch = reverse_vocab.get(ci, "<unk>")
        if ch != "<unk>":
            text += ch
            ids.append(ci)
            confidences.append(torch.max(prediction[len(text) - 1]).item())
        prev = ci
    conf = float(np.mean(confidences)) if confidences else 0.0
    return text, conf, ids


def recognize_batch(net: nn.Module, images: Sequence[np.ndarray], crop_w: int = 128,
                    return_logits: bool = False):
    """_recognize_crnn_batch, text_recognizer.py:114-140."""
    try:
        x = crnn_inputs(images, crop_w)
        with torch.no_grad():
            logits = net(x)
            preds = torch.softmax(logits, dim=2)
        res = []
        for p in preds:
            t, c, ids = decode_prediction(p)
            res.append({"text": t, "confidence": c, "ids": ids})
        return (res, logits) if return_logits else res
    except Exception:
        res = [{"text": "", "confidence": 0.0, "ids": []} for _ in images]
        return (res, None) if return_logits else res


# ----------------------------------------------------------------------------
# Pipeline per-frame body (pipeliine.py:104-139 / :143-172)
# ----------------------------------------------------------------------------
def process_frame(det_net, rec_net, frame: np.ndarray, threshold: float = 0.5,
                  det_h: int = 640, det_w: int = 640, crop_w: int = 128,
                  logit_bias: Optional[torch.Tensor] = None, per_crop: bool = True) -> List[Dict]:
    """detect -> crop original BGR frame -> recognise.  per_crop=True is what the
    pipeline does (batch-1 recognise per crop); False uses one batched call."""
    dets = detect(det_net, frame, threshold, det_h, det_w, logit_bias)
    crops, keep = [], []
    for d in dets:
        x1, y1, x2, y2 = d["bbox"]
        crop = frame[y1:y2, x1:x2]
        if crop.size == 0:                      # pipeliine.py:122-123
            continue
        crops.append(crop)
        keep.append(d)
    if per_crop:
        recs = [recognize_batch(rec_net, [c], crop_w)[0] for c in crops]
    else:
        recs = recognize_batch(rec_net, crops, crop_w) if crops else []
    out = []
    for d, r in zip(keep, recs):
        out.append({"bbox": d["bbox"], "text": r["text"], "detection_confidence": d["confidence"],
                    "recognition_confidence": r["confidence"], "polygon": d.get("polygon", []),
                    "ids": r["ids"]})
    return out


# ----------------------------------------------------------------------------
# Synthetic workloads
# ----------------------------------------------------------------------------
# synthetic_frames, planted_logit_bias and randomize_bn are input generators shared with bench.py's B200 arm: they live in
# video_text_detection_system_b200/synthetic.py (imported at the top) so that the product side never imports the oracle.
