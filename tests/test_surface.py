"""CPU: the reference's call surface (tests/test_models.py / test_integration.py of the reference, which patch
detect / recognize / forward) on our classes.  No GPU needed: every device call is behind a patched method."""
import asyncio
import json
import os
from unittest.mock import patch

import numpy as np
import pytest
import torch

from oracle import port
from video_text_detection_system_b200 import CRNN, DBNet, TextDetector, TextRecognizer, VideoTextPipeline
from video_text_detection_system_b200 import _lib


@pytest.fixture(scope="module")
def pipeline():
    return VideoTextPipeline(use_transformer_ocr=False, confidence_threshold=0.5, batch_size=16, backbone="resnet18",
                             pretrained=False)


def test_exports_match_reference_package():
    import video_text_detection_system_b200 as pkg
    for name in ("TextDetector", "DBNet", "TextRecognizer", "CRNN", "VideoTextPipeline"):   # app/ml/__init__.py:1-5
        assert hasattr(pkg, name)


def test_state_dict_layout_matches_reference():
    for bb in ("resnet18", "resnet50"):
        ours = DBNet(bb, pretrained=False).state_dict()
        ref = port.build_dbnet(bb).state_dict()           # torchvision + reference head/FPN naming
        assert list(ours) == list(ref)
        assert all(ours[k].shape == ref[k].shape for k in ref)
    assert list(CRNN(97).state_dict()) == list(port.build_crnn().state_dict())
    with pytest.raises(ValueError):
        DBNet("vgg16", pretrained=False)


def test_constructor_attributes(pipeline):
    d = TextDetector(backbone="resnet18", pretrained=False)
    assert isinstance(d.model, DBNet) and d.device is not None and callable(d.transform)
    r = TextRecognizer(use_transformer=False)
    assert r.use_transformer is False and isinstance(r.model, CRNN) and isinstance(r.vocab, dict)
    assert r.vocab == port.build_vocab() and len(r.vocab) == 97
    assert pipeline.batch_size == 16 and pipeline.confidence_threshold == 0.5
    for attr in ("detector", "recognizer", "video_processor", "image_processor", "executor"):
        assert hasattr(pipeline, attr)


def test_checkpoint_round_trip(tmp_path):
    sd = port.build_dbnet("resnet18", seed=5).state_dict()
    path = str(tmp_path / "det.pth")
    torch.save({"model_state_dict": sd}, path)
    d = TextDetector(path, backbone="resnet18", pretrained=False)
    assert all(torch.equal(d.model.state_dict()[k], sd[k]) for k in sd)
    with pytest.raises(Exception):
        TextDetector(str(tmp_path / "missing.pth"), backbone="resnet18", pretrained=False)
    crnn_sd = port.build_crnn(seed=2).state_dict()
    torch.save({"model_state_dict": crnn_sd}, str(tmp_path / "rec.pth"))
    r = TextRecognizer(str(tmp_path / "rec.pth"), use_transformer=False)
    assert torch.equal(r.model.state_dict()["classifier.weight"], crnn_sd["classifier.weight"])


def test_detect_never_raises():
    d = TextDetector(backbone="resnet18", pretrained=False)
    assert d.detect(None) == []
    assert d.detect(np.array([])) == []
    assert d.detect(np.zeros((10, 10), np.uint8)) == []
    assert d.detect(np.zeros((10, 10, 3), np.float32)) == []


def test_recognize_error_convention():
    r = TextRecognizer(use_transformer=False)
    out = r.recognize_batch([np.zeros((20, 40), np.uint8), np.zeros((20, 40), np.uint8)])   # 2-D crops fail in the reference too
    assert out == [{"text": "", "confidence": 0.0}] * 2
    assert r.recognize(np.zeros((0, 0, 3), np.uint8)) == {"text": "", "confidence": 0.0}


def test_process_single_frame_with_patched_models(pipeline):
    frame = np.random.randint(0, 255, (480, 640, 3), dtype=np.uint8)
    dets = [{"bbox": [10, 10, 100, 50], "confidence": 0.8, "polygon": [[10, 10], [100, 10], [100, 50], [10, 50]]}]
    with patch.object(pipeline.detector, "detect", return_value=dets), \
            patch.object(pipeline.recognizer, "recognize", return_value={"text": "sample", "confidence": 0.9}):
        res = pipeline.process_single_frame(frame)
    assert res == {"detections": [{"bbox": [10, 10, 100, 50], "text": "sample", "detection_confidence": 0.8,
                                   "recognition_confidence": 0.9}]}
    with patch.object(pipeline.detector, "detect", return_value=[]):
        assert pipeline.process_single_frame(frame) == {"detections": []}
    with patch.object(pipeline.detector, "detect", side_effect=RuntimeError("boom")):
        res = pipeline.process_single_frame(frame)
    assert res["detections"] == [] and "boom" in res["error"]
    # an empty crop is skipped (pipeliine.py:122-123)
    with patch.object(pipeline.detector, "detect", return_value=[{"bbox": [700, 10, 800, 50], "confidence": 0.8}]), \
            patch.object(pipeline.recognizer, "recognize", return_value={"text": "x", "confidence": 0.9}):
        assert pipeline.process_single_frame(frame) == {"detections": []}


def _make_video(path, n=30, size=(320, 240), fps=30.0):
    import cv2
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, size)
    for i in range(n):
        f = np.zeros((size[1], size[0], 3), np.uint8)
        cv2.putText(f, "HELLO %d" % i, (20, 120), cv2.FONT_HERSHEY_SIMPLEX, 1, (255, 255, 255), 2)
        vw.write(f)
    vw.release()


def test_process_video_with_patched_models(pipeline, tmp_path):
    video = str(tmp_path / "v.mp4")
    _make_video(video)
    if not os.path.exists(video) or os.path.getsize(video) == 0:
        pytest.skip("no mp4 encoder in this OpenCV build")
    dets = [{"bbox": [10, 10, 100, 50], "confidence": 0.8, "polygon": [[10, 10], [100, 10], [100, 50], [10, 50]]}]
    calls = []

    async def progress(p, done, total):
        calls.append((p, done, total))

    pipeline.batch_size = 4
    try:
        with patch.object(pipeline.detector, "detect", return_value=dets), \
                patch.object(pipeline.recognizer, "recognize", return_value={"text": " hi ", "confidence": 0.9}):
            res = asyncio.run(pipeline.process_video(video, str(tmp_path), progress))
    finally:
        pipeline.batch_size = 16
    assert res["status"] == "success"
    assert len(res["results"]) == 10                      # 30 fps source sampled at 10 fps (preprocessing.py:52)
    assert [r["frame_number"] for r in res["results"]] == list(range(10))
    d = res["results"][0]["detections"][0]
    assert set(d) == {"bbox", "text", "detection_confidence", "recognition_confidence", "polygon"}
    s = res["summary"]
    assert s["total_frames"] == 10 and s["frames_with_text"] == 10 and s["total_detections"] == 10
    assert s["unique_texts"] == 1 and s["detected_texts"] == ["hi"]
    assert s["avg_detection_confidence"] == pytest.approx(0.8) and s["fps_processed"] > 0
    assert calls and calls[-1][1] == 8
    json.dumps(res)                                       # stored in a JSON column: plain Python scalars only
    bad = asyncio.run(pipeline.process_video(str(tmp_path / "missing.mp4"), str(tmp_path)))
    assert bad["status"] in ("success", "failed") and bad["results"] == []


def test_summary_of_empty_results(pipeline):
    s = pipeline._generate_summary([], 0.0, 0)
    assert s["total_detections"] == 0 and s["avg_detection_confidence"] == 0.0 and s["fps_processed"] == 0


def test_records_to_detections_and_vocab():
    rec = np.zeros(3, _lib.RECORD_DTYPE)
    rec[0]["bbox"] = [1, 2, 30, 40]
    rec[0]["polygon"] = [1, 2, 30, 2, 30, 40, 1, 40]
    rec[0]["det_conf"], rec[0]["rec_conf"], rec[0]["len"] = 0.75, 0.5, 3
    rec[0]["ids"][:3] = [11, 12, 1]
    out = _lib.records_to_detections(rec, 1, with_text=True)
    assert out == [{"bbox": [1, 2, 30, 40], "confidence": 0.75, "polygon": [[1, 2], [30, 2], [30, 40], [1, 40]],
                    "ids": [11, 12, 1], "text": "ab0", "recognition_confidence": 0.5}]
    assert all(type(v) is int for v in out[0]["bbox"]) and type(out[0]["confidence"]) is float
    v = port.build_vocab()
    assert _lib.ids_to_text([v["0"], v["z"], v[" "], v["~"]]) == "0z ~"
    assert _lib.ids_to_text([0, 96, 200]) == ""


def test_frame_pointer_tables_bgr_and_nv12():
    """Host-side argument checking of the batch entry points (no GPU): BGR frames are HxWx3, NV12 frames (H*3/2)xW planes
    with even H and W; mixed sizes in one batch are refused."""
    import numpy as np
    import pytest
    from video_text_detection_system_b200 import _lib
    bgr = [np.zeros((36, 64, 3), np.uint8), np.zeros((36, 64, 3), np.uint8)]
    ptrs, h, w, pitch, keep = _lib.Engine._frame_ptrs(bgr)
    assert (h, w, pitch) == (36, 64, 192) and len(keep) == 2 and ptrs[0] == bgr[0].ctypes.data
    nv12 = [np.zeros((54, 64), np.uint8), np.zeros((54, 64), np.uint8)]
    ptrs, h, w, pitch, keep = _lib.Engine._frame_ptrs(nv12, _lib.VTD_PIX_NV12)
    assert (h, w, pitch) == (36, 64, 64)
    with pytest.raises(ValueError):
        _lib.Engine._frame_ptrs([np.zeros((54, 64, 3), np.uint8)], _lib.VTD_PIX_NV12)      # not a plane
    with pytest.raises(ValueError):
        _lib.Engine._frame_ptrs([np.zeros((53, 64), np.uint8)], _lib.VTD_PIX_NV12)         # rows not a multiple of 3
    with pytest.raises(ValueError):
        _lib.Engine._frame_ptrs([bgr[0], np.zeros((40, 64, 3), np.uint8)])                  # mixed sizes
    assert _lib.STAGE_NAMES[4] == "lstm0" and len(_lib.STAGE_NAMES) == 7


def test_records_to_regions_matches_per_record_definition():
    """The vectorised record -> dict conversion (one bytes.translate for a frame's texts) against the plain
    per-record definition, on random bytes: ids outside 1..95, lengths beyond the ids field, empty texts."""
    rng = np.random.default_rng(5)
    rec = np.zeros((4, 64), _lib.RECORD_DTYPE)
    rec["bbox"] = rng.integers(0, 1920, (4, 64, 4))
    rec["polygon"] = rng.integers(-5, 1312, (4, 64, 8))
    rec["det_conf"], rec["rec_conf"] = rng.random((4, 64)), rng.random((4, 64))
    rec["len"] = rng.integers(0, 45, (4, 64))
    rec["ids"] = rng.integers(0, 256, (4, 64, 36))
    rec["ids"][1] = rng.integers(1, 96, (64, 36))                 # a frame without any NUL after translation
    for i in range(4):
        count = (0, 64, 50, 1)[i]
        regions = _lib.records_to_regions(rec[i], count)
        dets = _lib.records_to_detections(rec[i], count, with_text=True)
        assert len(regions) == len(dets) == count
        for k, (r, d) in enumerate(zip(regions, dets)):
            n = min(int(rec[i, k]["len"]), 36)
            want = _lib.ids_to_text(rec[i, k]["ids"][:n].tolist())
            assert r["text"] == d["text"] == want
            assert r["bbox"] == d["bbox"] == rec[i, k]["bbox"].tolist()
            assert r["polygon"] == d["polygon"] == rec[i, k]["polygon"].reshape(4, 2).tolist()
            assert r["detection_confidence"] == d["confidence"] == float(rec[i, k]["det_conf"])
            assert r["recognition_confidence"] == d["recognition_confidence"] == float(rec[i, k]["rec_conf"])
            assert list(r) == ["bbox", "text", "detection_confidence", "recognition_confidence", "polygon"]
        json.dumps(regions)


def test_gc_paused_is_reentrant_and_restores_state():
    import gc
    import threading
    assert gc.isenabled()
    with _lib.gc_paused():
        assert not gc.isenabled()
        with _lib.gc_paused():
            assert not gc.isenabled()
        assert not gc.isenabled()
    assert gc.isenabled()
    ths = [threading.Thread(target=lambda: [_lib.gc_paused().__enter__() or _lib.gc_paused().__exit__() for _ in range(200)])
           for _ in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert gc.isenabled()
    gc.disable()
    try:
        with _lib.gc_paused():
            pass
        assert not gc.isenabled()                                  # a caller's own gc.disable() is respected
    finally:
        gc.enable()


def test_process_video_fused_path_keeps_frame_order_with_batches_in_flight(tmp_path):
    """The fused path of process_video (one detect_and_recognize call per batch, `inflight` batches on executor
    threads) with the device call replaced: results stay in frame order whatever order the batches finish in, the
    tail batch is processed, progress is reported per retired batch, and the collector is left as it was found."""
    import gc
    import time as _time
    video = str(tmp_path / "w.mp4")
    _make_video(video, n=66)                                       # 22 sampled frames: 5 batches of 4 + a tail of 2
    if not os.path.exists(video) or os.path.getsize(video) == 0:
        pytest.skip("no mp4 encoder in this OpenCV build")
    p = VideoTextPipeline(use_transformer_ocr=False, batch_size=4, backbone="resnet18", pretrained=False, inflight=3)
    seen = []

    def fake(frames, slot=0):
        k = len(seen)
        seen.append((len(frames), slot))
        _time.sleep(0.03 * ((7 - k) % 3))                          # later batches finish first
        return [[{"bbox": [1, 2, 30 + k, 40], "text": "b%d" % k, "detection_confidence": 0.5,
                  "recognition_confidence": 0.25, "polygon": [[1, 2], [30, 2], [30, 40], [1, 40]]}] for _ in frames]

    calls = []

    async def progress(pr, done, total):
        calls.append(done)

    frozen_before = gc.get_freeze_count()
    with patch.object(p, "detect_and_recognize", side_effect=fake):
        res = asyncio.run(p.process_video(video, str(tmp_path), progress))
    assert res["status"] == "success"
    assert [r["frame_number"] for r in res["results"]] == list(range(22))
    assert [r["detections"][0]["text"] for r in res["results"]] == ["b%d" % (i // 4) for i in range(22)]
    assert [n for n, _ in seen] == [4, 4, 4, 4, 4, 2] and {s for _, s in seen[:5]} == {0, 1, 2}
    assert calls == [4, 8, 12, 16, 20]                              # the reference reports per full batch (:63-65)
    assert res["summary"]["total_frames"] == 22 and res["summary"]["total_detections"] == 22
    assert gc.isenabled() and gc.get_freeze_count() <= frozen_before   # nothing of ours left in the permanent generation
    json.dumps(res)
