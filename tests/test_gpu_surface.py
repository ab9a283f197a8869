"""GPU tests of the reference-facing Python surface and of the less common entry points (NV12 ingest, ResNet50,
4K, crop chunking, concurrent callers).  The oracle (oracle/port.py) is the checker."""
import os
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T16 = os.environ.get("VTD_TEST_TIER16", "fp16")      # the shipped speed tier (IEEE half storage); bf16 = the build option


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    return p


@pytest.fixture(scope="module")
def E():
    from video_text_detection_system_b200 import _lib
    return _lib


def _nv12_from_bgr(bgr):
    import cv2
    h, w = bgr.shape[:2]
    i420 = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420)
    nv12 = np.empty((h * 3 // 2, w), np.uint8)
    nv12[:h] = i420[:h]
    u = i420[h:h + h // 4].reshape(h // 2, w // 2)
    v = i420[h + h // 4:].reshape(h // 2, w // 2)
    nv12[h:, 0::2] = u
    nv12[h:, 1::2] = v
    return nv12


def test_nv12_preprocess_matches_cv2_then_reference_transform(E, port):
    import ctypes as C
    import cv2
    h, w, dh, dw = 360, 640, 256, 448
    rng = np.random.default_rng(3)
    frames = [_nv12_from_bgr(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)),
              rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)]          # second: arbitrary YUV values
    eng = E.Engine(det_h=dh, det_w=dw, max_batch=2, max_src_h=h, max_src_w=w)
    ptrs = (C.c_void_p * 2)(*[f.ctypes.data for f in frames])
    eng._check(eng.lib.vtd_preprocess(eng.handle, C.cast(ptrs, C.POINTER(C.c_void_p)), 2, h, w, w, E.VTD_PIX_NV12, 0))
    x = eng.debug_tensor("input", 2)
    for i, f in enumerate(frames):
        bgr = cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12)
        assert np.array_equal(x[i], port.preprocess(bgr, dh, dw)[0].numpy())


def test_resnet50_speed_tier_maps_vs_oracle(E, port):
    net = port.build_dbnet("resnet50", seed=4)
    h, w = 192, 256
    x = np.random.default_rng(1).standard_normal((2, 3, h, w)).astype(np.float32)
    eng = E.Engine(backbone=50, det_h=h, det_w=w, max_batch=2, dtype=T16)
    eng.load_detector(net.state_dict())
    p, t = eng.dbnet_forward(x)
    with torch.no_grad():
        ref = port.dbnet_forward(net, torch.from_numpy(x), return_feats=True)
    p2 = eng.debug_tensor("p2", 2)
    rp2 = ref["p2"].numpy()
    ep, et = np.abs(p - ref["probability"].numpy()).max(), np.abs(t - ref["threshold"].numpy()).max()
    print("R50 %s 192x256: p2 rel %.2e, max |dprob| %.2e, max |dthresh| %.2e" % (T16, np.abs(p2 - rp2).max() / np.abs(rp2).max(), ep, et))
    over = float(np.mean(np.abs(p - ref["probability"].numpy()) > 1e-2))
    print("R50 %s 192x256: share of probability pixels over 1e-2: %.2e" % (T16, over))
    if T16 == "fp16":
        # The shipped tier, absolute bounds.  A random-init ResNet50 with randomised BN statistics has pre-sigmoid logits of
        # std ~30 (94 % of the pixels saturated): 53 layers of half rounding leave a relative logit error of ~2.6e-3
        # (asserted on p2 below), which at a zero crossing of a logit of that size moves the probability by up to 3.4e-2
        # (measured 3.3e-2 / 3.4e-2).  ResNet18's logits are O(1) and meet 1e-2 outright (tests/test_gpu_tiers.py).
        assert np.abs(p2 - rp2).max() <= 4e-3 * np.abs(rp2).max()
        assert ep <= 5e-2 and et <= 5e-2         # measured 3.3e-2 / 3.4e-2
        assert over <= 1.5e-2                    # measured 0.8 % of the pixels
    else:
        # bfloat16 build option: a 1 % error in a saturated logit that crosses zero moves the probability by more than 1e-2
        assert np.abs(p2 - rp2).max() <= 0.03 * np.abs(rp2).max()
        assert (np.abs(p - ref["probability"].numpy()) <= 1e-2).mean() >= 0.90
        assert (np.abs(t - ref["threshold"].numpy()) <= 1e-2).mean() >= 0.90


def test_resnet50_at_the_benched_detector_size_vs_oracle(E, port):
    """BASELINE configs[4]'s backbone (DBNet-ResNet50) at a size the CPU oracle can afford -- one 1080p frame at 736x1312,
    the detector size of configs[1..3] -- in the speed tier, with absolute tolerances, through preprocess + detect +
    box extraction."""
    net = port.build_dbnet("resnet50", seed=0)
    H, W, DH, DW = 1080, 1920, 736, 1312
    frames = port.synthetic_frames(1, H, W, seed=4)
    # +-1000: the random-init ResNet50's own logits reach +-150, so the planted plane has to be that much larger to decide the mask
    bias = port.planted_logit_bias(1, DH, DW, seed=5, boxes=50, inside=1000.0, outside=-1000.0)
    eng = E.Engine(backbone=50, det_h=DH, det_w=DW, max_batch=1, max_boxes=64, dtype=T16, max_src_h=H, max_src_w=W)
    eng.load_detector(net.state_dict())
    b = torch.from_numpy(bias).cuda()
    eng.preprocess(list(frames))
    eng.detect_maps(1, 0.5, b.data_ptr())
    p, t, m = eng.read_maps(1)
    eng.extract_boxes(1, H, W)
    rec, cnt = eng.read_records(1)
    assert eng.overflow() == 0
    with torch.no_grad():
        ref = port.dbnet_forward(net, port.preprocess(frames[0], DH, DW), torch.from_numpy(bias)[:, None])
    rp, rt = ref["probability"].numpy()[0, 0], ref["threshold"].numpy()[0, 0]
    ep, et = float(np.abs(p[0] - rp).max()), float(np.abs(t[0] - rt).max())
    print("R50 %s 736x1312: max |dprob| %.2e, max |dthresh| %.2e, threshold pixels over 1e-2: %.2e" % (T16, ep, et, float(np.mean(np.abs(t[0] - rt) > 1e-2))))
    if T16 == "fp16":
        assert ep <= 1e-2                      # the probability map is decided by the planted plane: exact to rounding
        # the threshold map is the net's own: measured max 4.5e-2, 0.84 % of the pixels over 1e-2 (see
        # test_resnet50_speed_tier_maps_vs_oracle for why a random-init ResNet50 cannot do better in 16 bits)
        assert et <= 7e-2 and np.mean(np.abs(t[0] - rt) > 1e-2) <= 1.5e-2
    want = sorted(tuple(d["bbox"]) for d in port.post_process(rp, W, H, 0.5, DH, DW))
    got = sorted(tuple(int(v) for v in r["bbox"]) for r in rec[0][:cnt[0]])
    assert len(want) >= 45 and got == want


def test_4k_resnet50_speed_tier_runs_and_agrees_with_fp32_tier(E, port):
    """BASELINE config 5 shape: 2160x3840 -> 2176x3840, DBNet-ResNet50, 16-bit tier.  Oracle-free (a CPU forward at this
    size takes minutes; the oracle comparison of this backbone is the 736x1312 test above): the two tiers of the
    library must agree within the 16-bit tolerance."""
    net = port.build_dbnet("resnet50", seed=0)
    frames = port.synthetic_frames(1, 2160, 3840, seed=9)
    outs = {}
    for dtype in ("fp32", T16):
        eng = E.Engine(backbone=50, det_h=2176, det_w=3840, max_batch=1, dtype=dtype, max_src_h=2160, max_src_w=3840)
        eng.load_detector(net.state_dict())
        eng.preprocess(list(frames))
        eng.detect_maps(1, 0.5)
        outs[dtype] = eng.read_maps(1)
        eng.close()
    assert np.isfinite(outs[T16][0]).all()
    ep, et = np.abs(outs["fp32"][0] - outs[T16][0]).max(), np.abs(outs["fp32"][1] - outs[T16][1]).max()
    print("4K R50 %s vs fp32 tier: max |dprob| %.2e, max |dthresh| %.2e" % (T16, ep, et))
    over = float(np.mean(np.abs(outs["fp32"][0] - outs[T16][0]) > 1e-2))
    print("4K R50 %s vs fp32 tier: share of probability pixels over 1e-2: %.2e" % (T16, over))
    if T16 == "fp16":
        # measured 1.0e-1 / 1.0e-1, 1.8 % of the pixels over 1e-2: 8.4 M pixels of saturated random-init logits (see
        # test_resnet50_speed_tier_maps_vs_oracle); finite everywhere, i.e. no half overflow at this depth and size
        assert ep <= 0.15 and et <= 0.15 and over <= 3e-2
    else:
        assert (np.abs(outs["fp32"][0] - outs[T16][0]) <= 1e-2).mean() >= 0.90
        assert (np.abs(outs["fp32"][1] - outs[T16][1]) <= 1e-2).mean() >= 0.90


def test_more_crops_than_one_chunk(E, port):
    net = port.build_crnn(seed=0)
    eng = E.Engine(det_h=32, det_w=32, crop_w=128, max_batch=1, max_boxes=64, max_src_h=32, max_src_w=32)   # chunk = 64
    eng.load_recognizer(net.state_dict())
    rng = np.random.default_rng(2)
    crops = [rng.integers(0, 256, (rng.integers(12, 40), rng.integers(20, 150), 3), dtype=np.uint8) for _ in range(150)]
    ids, lens, conf, logits = eng.recognize_crops(crops, want_logits=True)
    ref, ref_logits = port.recognize_batch(net, crops, return_logits=True)
    assert np.abs(logits - ref_logits.numpy()).max() <= 5e-3
    agree = sum(ids[i, :lens[i]].tolist() == ref[i]["ids"] for i in range(len(crops)))
    assert agree >= 0.95 * len(crops)          # 1-LSB resize differences may flip a near-tie argmax


def test_text_detector_detect_vs_oracle_and_threads(port):
    from video_text_detection_system_b200 import TextDetector
    sd = port.build_dbnet("resnet18", seed=0).state_dict()
    D = TextDetector(backbone="resnet18", pretrained=False, det_size=(320, 480), dtype="fp32")     # exact boxes: parity tier
    D.model.load_state_dict(sd)
    net = port.build_dbnet("resnet18", seed=0)
    import cv2
    frames = []
    for i in range(4):
        f = np.zeros((360, 540, 3), np.uint8)
        cv2.putText(f, "FRAME %d" % i, (30, 120 + 40 * i), cv2.FONT_HERSHEY_SIMPLEX, 2, (255, 255, 255), 5)
        frames.append(f)
    want = [port.detect(net, f, 0.5, 320, 480) for f in frames]
    got = [None] * 4

    def work(i):
        got[i] = D.detect(frames[i], 0.5)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]     # the reference's 4 executor threads
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for g, w_ in zip(got, want):
        assert sorted(tuple(d["bbox"]) for d in g) == sorted(tuple(d["bbox"]) for d in w_)
        for d in g:
            assert isinstance(d["confidence"], float) and all(isinstance(v, int) for v in d["bbox"])
    # patched forward (reference tests/test_models.py:30-37): arbitrary map size, post-process only
    from unittest.mock import patch
    with patch.object(D.model, "forward", return_value={"probability": torch.rand(1, 1, 160, 160),
                                                        "threshold": torch.rand(1, 1, 160, 160)}):
        out = D.detect(frames[0])
    assert isinstance(out, list)
    pm = np.random.default_rng(0).random((160, 160))
    got_pp = D._post_process(pm, 640, 480, 0.5)
    want_pp = port.post_process(pm.astype(np.float32), 640, 480, 0.5, 320, 480)
    assert sorted(tuple(d["bbox"]) for d in got_pp) == sorted(tuple(d["bbox"]) for d in want_pp)


def test_pipeline_fused_path_vs_oracle(port):
    from video_text_detection_system_b200 import VideoTextPipeline
    # exact box sets and text: the fp32 parity tier (the default speed tier is covered by tests/test_gpu_tiers.py)
    P = VideoTextPipeline(use_transformer_ocr=False, backbone="resnet18", pretrained=False, det_size=(256, 1280), dtype="fp32")
    det, rec = port.build_dbnet("resnet18", seed=0), port.build_crnn(seed=0)
    P.detector.model.load_state_dict(det.state_dict())
    P.recognizer.model.load_state_dict(rec.state_dict())
    import cv2
    frames = []
    for i in range(3):
        f = np.full((288, 1440, 3), 30, np.uint8)
        cv2.putText(f, "HELLO WORLD %d" % i, (40, 200), cv2.FONT_HERSHEY_SIMPLEX, 4, (255, 255, 255), 9)
        frames.append(f)
    got = P.detect_and_recognize(frames)
    for f, regions in zip(frames, got):
        want = port.process_frame(det, rec, f, 0.5, 256, 1280, 128, per_crop=False)
        assert sorted(tuple(r["bbox"]) for r in regions) == sorted(tuple(r["bbox"]) for r in want)
        for r in regions:
            assert set(r) == {"bbox", "text", "detection_confidence", "recognition_confidence", "polygon"}
    single = P.process_single_frame(frames[0])
    assert [r["bbox"] for r in single["detections"]] == [r["bbox"] for r in got[0]]
    assert all("polygon" not in r for r in single["detections"])
    # recogniser surface
    crop = frames[0][100:220, 40:700]
    r1 = P.recognizer.recognize(crop)
    w1 = port.recognize_batch(rec, [crop])[0]
    assert r1["text"] == w1["text"] and r1["confidence"] == pytest.approx(w1["confidence"], abs=2e-3)


def test_nv12_run_batch_equals_bgr_run_batch(E, port):
    """Decoder-surface ingest: the whole path on NV12 frames -- preprocess AND the crop gather read
    the NV12 planes directly -- gives the very records it gives on cv2.cvtColor(COLOR_YUV2BGR_NV12) of those frames."""
    import cv2
    h, w, dh, dw, n = 180, 360, 160, 320, 2
    det, rec = port.build_dbnet("resnet18", seed=0), port.build_crnn(seed=0)
    rng = np.random.default_rng(11)
    nv12 = [_nv12_from_bgr(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)),
            rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)]
    bgr = [cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12) for f in nv12]
    bias = torch.from_numpy(port.planted_logit_bias(n, dh, dw, seed=2, boxes=4)).cuda()
    out = {}
    for name, frames, pix in (("nv12", nv12, E.VTD_PIX_NV12), ("bgr", bgr, E.VTD_PIX_BGR)):
        eng = E.Engine(det_h=dh, det_w=dw, max_batch=n, max_boxes=32, max_src_h=h, max_src_w=w)
        eng.load_detector(det.state_dict())
        eng.load_recognizer(rec.state_dict())
        out[name] = eng.run_batch(frames, thr=0.5, recognize=True, logit_bias_dev=bias.data_ptr(), pixfmt=pix)
        eng.close()
    assert out["bgr"][1].sum() > 0
    assert np.array_equal(out["nv12"][1], out["bgr"][1])
    for i in range(n):
        k = out["bgr"][1][i]
        assert out["nv12"][0][i, :k].tobytes() == out["bgr"][0][i, :k].tobytes()
