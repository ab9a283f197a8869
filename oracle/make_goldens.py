"""Mint tests/golden/*.npz from the REFERENCE's own code.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):  python -m oracle.make_goldens
Every array written here is an output of the reference's source files (loaded by
oracle/reference_loader.py), not of oracle/port.py.  Weights are not stored: they
are the oracle builders' seeded initialisation copied into the reference modules
via load_state_dict (key layout identical); a checksum of
the state dict is stored so RNG drift is detected instead of mis-read as a
parity failure.
"""
from __future__ import annotations

import json
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import port, reference_loader as RL  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sd_checksum(sd) -> float:
    return float(sum(v.double().abs().sum().item() for v in sd.values() if v.dtype.is_floating_point))


def structured_frame(h=480, w=640, text="TEST TEXT"):
    """tests/test_models.py:15-19 style frame plus gradients so every channel varies."""
    f = np.zeros((h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    f[..., 0] = (xx * 255 // max(w - 1, 1)).astype(np.uint8)
    f[..., 1] = (yy * 255 // max(h - 1, 1)).astype(np.uint8)
    f[..., 2] = ((xx + yy) % 256).astype(np.uint8)
    cv2.putText(f, text, (50, 100), cv2.FONT_HERSHEY_SIMPLEX, 1, (255, 255, 255), 2)
    return f


def planted_map(rects, h=640, w=640, lo=0.05):
    pm = np.full((h, w), lo, np.float32)
    for (cx, cy, rw, rh, ang, val) in rects:
        m = np.zeros((h, w), np.uint8)
        pts = cv2.boxPoints(((float(cx), float(cy)), (float(rw), float(rh)), float(ang)))
        cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
        pm[m > 0] = val
    return pm


def postprocess_cases():
    """Hand-built maps for the semantics of SURVEY.md Appendix B.3."""
    cases = {}
    pm = np.zeros((640, 640), np.float32)
    pm[100:140, 200:400] = 0.9
    pm[300:330, 50:90] = 0.8
    cases["two_rects"] = (pm, 1920, 1080, 0.5)
    pm = np.zeros((640, 640), np.float32)          # ring + island: island is not external
    pm[100:200, 100:300] = 0.9
    pm[120:180, 120:280] = 0.1
    pm[135:165, 150:250] = 0.95
    cases["ring_island"] = (pm, 1280, 720, 0.5)
    pm = np.zeros((640, 640), np.float32)          # area filter: 11x11 -> 100 kept?, 10x11 -> 90 dropped
    pm[50:61, 50:61] = 0.9
    pm[50:60, 100:111] = 0.9
    pm[200:230, 200:260] = 0.7
    cases["area_filter"] = (pm, 640, 640, 0.5)
    pm = np.zeros((640, 640), np.float32)          # diagonal touch = one 8-connected component
    pm[300:320, 300:340] = 0.9
    pm[320:345, 340:390] = 0.9
    cases["diag_touch"] = (pm, 1920, 1080, 0.5)
    cases["rotated"] = (planted_map([(320, 320, 200, 60, 30, 0.8), (150, 500, 120, 40, -20, 0.9),
                                     (500, 120, 90, 30, 75, 0.75)]), 1920, 1080, 0.5)
    pm = np.zeros((640, 640), np.float32)          # border-touching blobs
    pm[0:30, 0:120] = 0.9
    pm[600:640, 560:640] = 0.85
    pm[300:340, 0:50] = 0.6
    cases["border"] = (pm, 1920, 1080, 0.5)
    pm = np.full((640, 640), 0.5, np.float32)      # strict '>' : 0.5 is NOT foreground
    pm[400:440, 100:300] = np.float32(0.5000001)
    cases["strict_gt"] = (pm, 800, 600, 0.5)
    pm = np.zeros((640, 640), np.float32)          # concave (L and U shapes), thin line
    pm[100:200, 100:130] = 0.9
    pm[170:200, 100:260] = 0.9
    pm[300:400, 300:320] = 0.8
    pm[300:400, 400:420] = 0.8
    pm[380:400, 300:420] = 0.8
    pm[500:502, 100:400] = 0.9
    cases["concave"] = (pm, 1920, 1080, 0.5)
    rng = np.random.default_rng(7)
    cases["grid50"] = (1.0 / (1.0 + np.exp(-port.planted_logit_bias(1, 640, 640, seed=3)[0])), 1920, 1080, 0.5)
    pm = rng.random((160, 160)).astype(np.float32)   # tests/test_models.py:51 shape (map != 640)
    cases["random160"] = (pm, 640, 480, 0.5)
    blobs = (cv2.GaussianBlur(rng.random((640, 640)).astype(np.float32), (0, 0), 6) - 0.5) * 8 + 0.5
    cases["blobs"] = (np.clip(blobs, 0, 1).astype(np.float32), 1920, 1080, 0.5)
    return cases


def main():
    os.makedirs(OUT, exist_ok=True)
    assert RL.available(), "reference tree not mounted"
    det_mod, rec_mod = RL.modules()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)

    # ---- detector network: reference FPN/DBHead classes, repaired wiring -------------
    for bb, (h, w) in (("resnet18", (64, 96)), ("resnet50", (64, 64))):
        o = port.build_dbnet(bb, seed=0)
        ref = RL.reference_dbnet(bb, o.state_dict())
        x = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 3, h, w)).astype(np.float32))
        with torch.no_grad():
            r = port.dbnet_forward(ref, x, return_feats=True)
        np.savez_compressed(os.path.join(OUT, f"dbnet_{bb}.npz"), x=x.numpy(),
                            probability=r["probability"].numpy(), threshold=r["threshold"].numpy(),
                            c2_s=r["c2"].numpy()[:, ::4, ::2, ::2].copy(), c5=r["c5"].numpy(),
                            p2_s=r["p2"].numpy()[:, ::8, ::2, ::2].copy(),
                            seed=0, sd_checksum=sd_checksum(o.state_dict()))

    # ---- preprocess: the reference TextDetector.transform (640x640) -------------------
    o18 = port.build_dbnet("resnet18", seed=0)
    ref18 = RL.reference_dbnet("resnet18", o18.state_dict())
    D = RL.reference_detector(ref18)
    frames = {"structured": structured_frame(),
              "small_random": np.random.default_rng(2).integers(0, 256, (24, 32, 3), dtype=np.uint8),
              "hd_gradient": structured_frame(540, 960, "HELLO WORLD")}
    pre = {}
    for k, f in frames.items():
        t = D.transform(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)).numpy()
        mean = np.asarray(port.IMAGENET_MEAN, np.float32)[:, None, None]
        std = np.asarray(port.IMAGENET_STD, np.float32)[:, None, None]
        u8 = np.rint((t * std + mean) * 255.0).astype(np.uint8)          # exact inverse of ToTensor+Normalize
        pre[f"{k}_frame"] = f
        pre[f"{k}_resized_rgb_u8"] = u8
        pre[f"{k}_tensor_sample"] = t[:, ::37, ::41].copy()              # fp32 spot samples
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **pre)

    # ---- post-process: reference TextDetector._post_process ---------------------------
    pp = {}
    for k, (pm, ow, oh, thr) in postprocess_cases().items():
        dets = D._post_process(pm, ow, oh, thr)
        pp[f"{k}_map"] = pm
        pp[f"{k}_args"] = np.asarray([ow, oh, thr], np.float64)
        pp[f"{k}_dets"] = np.frombuffer(json.dumps(dets).encode(), np.uint8)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **pp)

    # ---- CTC decode: reference TextRecognizer._decode_prediction ----------------------
    ocr = port.build_crnn(seed=0)
    R = RL.reference_recognizer(ocr.state_dict())
    rng = np.random.default_rng(5)
    seqs = [[30, 0, 30, 31], [30, 30, 31, 31, 0], [96, 30, 96, 30], [30, 31, 30], [0] * 8,
            [96, 96, 0, 96], [1], [0, 0, 95, 95, 0, 95, 2], list(rng.integers(0, 97, 31)),
            list(rng.integers(0, 4, 31)), list(rng.integers(0, 97, 24))]
    ctc = {}
    for i, s in enumerate(seqs):
        T = len(s)
        logits = rng.standard_normal((T, 97)).astype(np.float32)
        logits[np.arange(T), s] += 6.0
        p = torch.softmax(torch.from_numpy(logits), dim=1)
        text, conf = R._decode_prediction(p)
        ctc[f"{i}_logits"] = logits
        ctc[f"{i}_text"] = np.frombuffer(text.encode(), np.uint8)
        ctc[f"{i}_conf"] = np.float64(conf)
    # ties -> lowest index (argmax), all-equal rows
    p = torch.full((5, 97), 1.0 / 97)
    text, conf = R._decode_prediction(p)
    ctc["tie_probs"] = p.numpy()
    ctc["tie_text"] = np.frombuffer(text.encode(), np.uint8)
    ctc["tie_conf"] = np.float64(conf)
    ctc["n"] = len(seqs)
    np.savez_compressed(os.path.join(OUT, "ctc.npz"), **ctc)

    # ---- CRNN: reference CRNN + _recognize_crnn_batch ---------------------------------
    crops = [rng.integers(0, 256, (40, 200, 3), dtype=np.uint8),
             structured_frame(64, 256, "abc")[:, :, :],
             rng.integers(0, 256, (17, 33, 3), dtype=np.uint8),
             structured_frame(32, 128, "Zq9")]
    res = R.recognize_batch(crops)
    with torch.no_grad():
        x = port.crnn_inputs(crops)
        logits = R.model(x).numpy()
    cr = {f"crop{i}": c for i, c in enumerate(crops)}
    cr.update(n=len(crops), logits=logits, inputs=x.numpy(),
              results=np.frombuffer(json.dumps(res).encode(), np.uint8),
              sd_checksum=sd_checksum(ocr.state_dict()))
    np.savez_compressed(os.path.join(OUT, "crnn.npz"), **cr)

    # ---- pipeline: reference process_single_frame body (pipeliine.py:143-172) with the
    #      detector's model.forward patched to a planted map, as tests/test_models.py does --
    frame = structured_frame(480, 640, "PIPELINE")
    pm = planted_map([(200, 150, 220, 60, 0, 0.9), (420, 400, 180, 50, 10, 0.8), (120, 520, 150, 44, -12, 0.85)])
    D.model.forward = lambda x: {"probability": torch.from_numpy(pm)[None, None],
                                 "threshold": torch.zeros(1, 1, 640, 640)}
    dets = D.detect(frame, 0.5)
    regions = []
    for d in dets:                                   # pipeliine.py:152-166
        x1, y1, x2, y2 = d["bbox"]
        crop = frame[y1:y2, x1:x2]
        if crop.size == 0:
            continue
        t = R.recognize(crop)
        regions.append({"bbox": d["bbox"], "text": t["text"], "detection_confidence": d["confidence"],
                        "recognition_confidence": t["confidence"], "polygon": d["polygon"]})
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), frame=frame, planted_map=pm,
                        regions=np.frombuffer(json.dumps(regions).encode(), np.uint8))
    print("goldens written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
