"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes -> libvtd_b200.so), against the CPU oracle
(oracle/port.py) and the golden vectors minted from the reference's own code (tests/golden/*.npz).

Tolerances (BASELINE.json north_star): probability/threshold maps <= 1e-3 abs in the fp32 tier, <= 1e-2 in
the 16-bit speed tier (IEEE half storage in the shipped library); identical masks away from threshold ties; box sets identical as integer sets (IoU >= 0.99 bar);
bit-exact CTC token ids for the same logits; preprocess bit-exact; crop resize within 1 LSB.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden

pytestmark = pytest.mark.gpu
# The speed tier the tests run: "fp16" = the shipped library (IEEE half storage).  VTD_TEST_TIER16=bf16 runs the same tests
# against libvtd_b200_bf16.so (bfloat16 storage), whose looser 640x640 figures are stated in that test.
T16 = os.environ.get("VTD_TEST_TIER16", "fp16")
STORAGE_16 = torch.float16 if T16 == "fp16" else torch.bfloat16


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    return p


@pytest.fixture(scope="module")
def E():
    from video_text_detection_system_b200 import _lib
    return _lib


def iou(a, b):
    x1, y1, x2, y2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    inter = max(0, x2 - x1) * max(0, y2 - y1)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def match_boxes(mine, ref, thr=0.99):
    """Every reference box has a distinct partner at IoU >= thr and vice versa."""
    assert len(mine) == len(ref), (len(mine), len(ref))
    used = set()
    for r in ref:
        best, bi = -1, -1
        for i, m in enumerate(mine):
            if i in used:
                continue
            v = iou(m["bbox"], r["bbox"])
            if v > best:
                best, bi = v, i
        assert best >= thr, (r, best)
        used.add(bi)


# ---------------------------------------------------------------- stage 1: preprocess
def test_preprocess_bit_exact_golden(E):
    g = load_golden("preprocess")
    eng = E.Engine(det_h=640, det_w=640, max_batch=1, max_src_h=1080, max_src_w=1920)
    mean = np.asarray([0.485, 0.456, 0.406], np.float32)[:, None, None]
    std = np.asarray([0.229, 0.224, 0.225], np.float32)[:, None, None]
    for k in ("structured", "small_random", "hd_gradient"):
        frame = g[k + "_frame"]
        eng.preprocess([frame])
        x = eng.debug_tensor("input", 1)[0]
        u8 = g[k + "_resized_rgb_u8"]
        want = (u8.astype(np.float32) / np.float32(255.0) - mean) / std
        assert np.array_equal(x, want), k
        assert np.array_equal(x[:, ::37, ::41], g[k + "_tensor_sample"]), k


@pytest.mark.parametrize("src,det", [((1080, 1920), (736, 1312)), ((480, 640), (640, 640)), ((300, 500), (320, 352))])
def test_preprocess_vs_oracle(E, port, src, det):
    frames = port.synthetic_frames(2, src[0], src[1], seed=3)
    eng = E.Engine(det_h=det[0], det_w=det[1], max_batch=2, max_src_h=src[0], max_src_w=src[1])
    eng.preprocess(list(frames))
    x = eng.debug_tensor("input", 2)
    for i in range(2):
        want = port.preprocess(frames[i], det[0], det[1])[0].numpy()      # PIL + torchvision, the reference's own calls
        assert np.array_equal(x[i], want)
    # 16-bit tier: same integers, rounded once
    engb = E.Engine(det_h=det[0], det_w=det[1], max_batch=2, max_src_h=src[0], max_src_w=src[1], dtype=T16)
    engb.preprocess(list(frames))
    xb = engb.debug_tensor("input", 2)
    assert np.array_equal(xb, torch.from_numpy(x).to(STORAGE_16).float().numpy())


# ---------------------------------------------------------------- stages 2+3: DBNet + fused head
@pytest.mark.parametrize("bb", ["resnet18", "resnet50"])
def test_dbnet_golden_fp32(E, port, bb):
    g = load_golden("dbnet_" + bb)
    net = port.build_dbnet(bb, seed=0)
    x = g["x"]
    eng = E.Engine(backbone=18 if bb == "resnet18" else 50, det_h=x.shape[2], det_w=x.shape[3], max_batch=2)
    eng.load_detector(net.state_dict())
    p, t = eng.dbnet_forward(x)
    assert np.abs(p - g["probability"]).max() <= 1e-3
    assert np.abs(t - g["threshold"]).max() <= 1e-3
    c5 = eng.debug_tensor("c5", 2)
    assert np.abs(c5 - g["c5"]).max() <= 2e-3 * max(1.0, np.abs(g["c5"]).max())
    p2 = eng.debug_tensor("p2", 2)[:, ::8, ::2, ::2]
    assert np.abs(p2 - g["p2_s"]).max() <= 2e-3 * max(1.0, np.abs(g["p2_s"]).max())


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-3), (T16, 1e-2)])
def test_dbnet_maps_vs_oracle(E, port, dtype, tol):
    net = port.build_dbnet("resnet18", seed=1)
    h, w = 160, 224
    frames = port.synthetic_frames(3, 270, 480, seed=5)
    eng = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=3, dtype=dtype, max_src_h=270, max_src_w=480)
    eng.load_detector(net.state_dict())
    eng.preprocess(list(frames))
    eng.detect_maps(3, 0.5)
    p, t, m = eng.read_maps(3)
    x = torch.cat([port.preprocess(f, h, w) for f in frames])
    with torch.no_grad():
        ref = port.dbnet_forward(net, x)
    rp, rt = ref["probability"].numpy()[:, 0], ref["threshold"].numpy()[:, 0]
    assert np.abs(p - rp).max() <= tol, np.abs(p - rp).max()
    assert np.abs(t - rt).max() <= tol, np.abs(t - rt).max()
    # the mask is exactly `own prob > thr`, and agrees with the reference away from threshold ties
    assert np.array_equal(m, (p > 0.5).astype(np.uint8))
    far = np.abs(rp - 0.5) > tol
    assert np.array_equal(m[far], (rp > 0.5).astype(np.uint8)[far])


def test_conv_tcgen05_matches_cuda_core_path(E, port):
    """16-bit tier: tcgen05 implicit GEMM vs the fp32 FFMA tier on the same weights, layer by layer."""
    net = port.build_dbnet("resnet18", seed=2)
    h, w = 128, 192
    x = np.random.default_rng(0).standard_normal((2, 3, h, w)).astype(np.float32)
    e32 = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=2, dtype="fp32")
    e16 = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=2, dtype=T16, fuse_head=False)     # keeps the "head" feature map
    for e in (e32, e16):
        e.load_detector(net.state_dict())
    p32, t32 = e32.dbnet_forward(x)
    p16, t16 = e16.dbnet_forward(x)
    for name in ("c2", "c3", "c4", "c5", "p2_in", "p2", "head"):
        a, b = e32.debug_tensor(name, 2), e16.debug_tensor(name, 2)
        scale = np.abs(a).max()
        err = np.abs(a - b).max() / scale
        assert err < 0.05, (name, err)
    assert np.abs(p32 - p16).max() <= 1e-2
    assert np.abs(t32 - t16).max() <= 1e-2


def test_halo_and_direct_window_layers_vs_oracle(E, port):
    """16-bit tier at a size where the one-patch-per-tile ("halo") convolutions and the direct-window stem are the ones
    that run: 352x1024 -> stem output 176x512 (four full 128-pixel tiles per row), layer1 / FPN maps 88x256 (8x16 tiles
    with a padded last tile row: 88 = 5.5 x 16), layer2 44x128.  Oracle: the reference's modules on the same input."""
    net = port.build_dbnet("resnet18", seed=3)
    h, w = 352, 1024
    frames = port.synthetic_frames(2, 396, 1152, seed=6)
    x = torch.cat([port.preprocess(f, h, w) for f in frames])
    eng = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=2, dtype=T16)
    eng.load_detector(net.state_dict())
    p, t = eng.dbnet_forward(x.numpy())
    with torch.no_grad():
        ref = port.dbnet_forward(net, x, return_feats=True)
    for name in ("c2", "c3", "p2"):
        got, want = eng.debug_tensor(name, 2), ref[name].numpy()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 0.03 * np.abs(want).max(), name
    assert np.abs(p - ref["probability"].numpy()).max() <= 1e-2
    assert np.abs(t - ref["threshold"].numpy()).max() <= 1e-2


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-3), ("fp16", 1e-2), ("bf16", 1e-2)])
def test_config1_640x640_maps_and_boxes_vs_oracle(E, port, dtype, tol):
    """BASELINE configs[0] shape -- the reference's own 640x640 detector input (text_detector.py:101) -- through the
    whole detect path: maps within the tier's tolerance, mask exact, boxes as the reference's post-process finds them on
    the library's own probability map."""
    net = port.build_dbnet("resnet18", seed=0)
    frames = port.synthetic_frames(2, 640, 640, seed=0)
    eng = E.Engine(backbone=18, det_h=640, det_w=640, max_batch=2, dtype=dtype, max_src_h=640, max_src_w=640, max_boxes=64)
    eng.load_detector(net.state_dict())
    bias = port.planted_logit_bias(2, 640, 640, seed=3, boxes=12)
    b = torch.from_numpy(bias).cuda()
    eng.preprocess(list(frames))
    eng.detect_maps(2, 0.5)
    p, t, m = eng.read_maps(2)
    x = torch.cat([port.preprocess(f, 640, 640) for f in frames])
    with torch.no_grad():
        ref = port.dbnet_forward(net, x)
    for got, want in ((p, ref["probability"].numpy()[:, 0]), (t, ref["threshold"].numpy()[:, 0])):
        err = np.abs(got - want)
        print("%s 640x640: max %.2e, pixels over %g: %d of %d" % (dtype, err.max(), tol, int((err > tol).sum()), err.size))
        if dtype != "bf16":
            # fp32 tier <= 1e-3; the shipped speed tier (IEEE half storage) meets north_star's 1e-2 outright
            # (measured 2.0e-3; CPU emulation of the rounding points: profiles/r01_bf16_error_budget.md)
            assert err.max() <= tol, err.max()
        else:
            # libvtd_b200_bf16.so, the bfloat16-storage BUILD OPTION (not the shipped tier): 819 200 pixels per map, measured
            # 0.11 % of the probability pixels over 1e-2 (max 1.6e-2), 0.33 % of the threshold pixels (max 2.0e-2) --
            # eight mantissa bits after each of ~25 layers.  Stated, bounded, and the reason half is what ships.
            assert (err > tol).mean() <= 5e-3 and err.max() <= 3e-2, (err.max(), int((err > tol).sum()))
    assert np.array_equal(m, (p > 0.5).astype(np.uint8))
    # boxes: with the planted plane (random-init maps hold no component of 100 px^2)
    eng.detect_maps(2, 0.5, b.data_ptr())
    p, t, m = eng.read_maps(2)
    assert np.array_equal(m, (p > 0.5).astype(np.uint8))
    eng.extract_boxes(2, 640, 640)
    rec, cnt = eng.read_records(2)
    for i in range(2):
        want = sorted(tuple(d["bbox"]) for d in port.post_process(p[i], 640, 640, 0.5, 640, 640))
        got = sorted(tuple(int(v) for v in r["bbox"]) for r in rec[i][:cnt[i]])
        assert got == want and len(got) >= 1


def test_logit_bias_plants_boxes(E, port):
    net = port.build_dbnet("resnet18", seed=0)
    h, w = 256, 1280
    bias = port.planted_logit_bias(1, h, w, seed=4, boxes=10)
    frames = port.synthetic_frames(1, 288, 1440, seed=2)
    eng = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=1, max_src_h=288, max_src_w=1440)
    eng.load_detector(net.state_dict())
    eng.preprocess(list(frames))
    b = torch.from_numpy(bias).cuda()
    eng.detect_maps(1, 0.5, b.data_ptr())
    p, _, m = eng.read_maps(1)
    x = port.preprocess(frames[0], h, w)
    with torch.no_grad():
        rp = port.dbnet_forward(net, x, torch.from_numpy(bias)[:, None])["probability"].numpy()[0, 0]
    assert np.abs(p[0] - rp).max() <= 1e-3
    eng.extract_boxes(1, 288, 1440)
    rec, cnt = eng.read_records(1)
    mine = E.records_to_detections(rec[0], int(cnt[0]), False)
    ref = port.post_process(rp, 1440, 288, 0.5, h, w)
    assert len(ref) >= 8
    match_boxes(mine, ref)


# ---------------------------------------------------------------- stage 4a: box extraction
def _pp_cases():
    g = load_golden("postprocess")
    return sorted({k[:-4] for k in g.files if k.endswith("_map")})


@pytest.mark.parametrize("case", _pp_cases())
def test_postprocess_golden(E, case):
    g = load_golden("postprocess")
    pm = g[case + "_map"]
    ow, oh, thr = g[case + "_args"]
    ref = golden_json(g[case + "_dets"])
    eng = E.Engine(det_h=640, det_w=640, max_batch=1, max_boxes=1024)
    rec = eng.postprocess_map(pm, int(ow), int(oh), float(thr), clip_h=640, clip_w=640)
    mine = E.records_to_detections(rec, len(rec), False)
    key = lambda d: tuple(d["bbox"])
    assert sorted(map(key, mine)) == sorted(map(key, ref))
    rd = {key(d): d for d in ref}
    for d in mine:
        r = rd[key(d)]
        assert d["polygon"] == r["polygon"]
        assert d["confidence"] == pytest.approx(r["confidence"], rel=1e-5, abs=1e-6, nan_ok=True)


def test_postprocess_random_maps_vs_oracle(E, port):
    import cv2
    rng = np.random.default_rng(17)
    eng = E.Engine(det_h=736, det_w=1312, max_batch=1, max_boxes=1024)
    for i in range(6):
        f = cv2.GaussianBlur(rng.random((736, 1312)).astype(np.float32), (0, 0), 3 + 2 * i)
        pm = np.clip((f - 0.5) * (6 + 3 * i) + 0.5, 0, 1).astype(np.float32)
        ref = port.post_process(pm, 1920, 1080, 0.5, 736, 1312)
        rec = eng.postprocess_map(pm, 1920, 1080, 0.5, clip_h=736, clip_w=1312)
        mine = E.records_to_detections(rec, len(rec), False)
        assert sorted(tuple(d["bbox"]) for d in mine) == sorted(tuple(d["bbox"]) for d in ref), i
        assert sorted(map(json.dumps, (d["polygon"] for d in mine))) == sorted(map(json.dumps, (d["polygon"] for d in ref)))
    # pathological: the 4x4 block texture random-init weights give (~1e5 components, none >= 100 px^2)
    tex = np.kron(rng.random((184, 328)) > 0.5, np.ones((4, 4))).astype(np.float32)
    ref = port.post_process(tex, 1920, 1080, 0.5, 736, 1312)
    rec = eng.postprocess_map(tex, 1920, 1080, 0.5, clip_h=736, clip_w=1312)
    assert sorted(tuple(int(v) for v in r["bbox"]) for r in rec) == sorted(tuple(d["bbox"]) for d in ref)
    # empty and full maps
    assert len(eng.postprocess_map(np.zeros((736, 1312), np.float32), 1920, 1080, 0.5)) == 0
    full = eng.postprocess_map(np.ones((736, 1312), np.float32), 1920, 1080, 0.5, clip_h=736, clip_w=1312)
    reff = port.post_process(np.ones((736, 1312), np.float32), 1920, 1080, 0.5, 736, 1312)
    assert [tuple(int(v) for v in r["bbox"]) for r in full] == [tuple(d["bbox"]) for d in reff]


# ---------------------------------------------------------------- stage 4b: crops, CRNN, CTC
def test_crop_resize_within_one_lsb(E, port):
    import cv2
    rng = np.random.default_rng(8)
    eng = E.Engine(det_h=32, det_w=32, crop_w=128, max_batch=1, max_boxes=64, max_src_h=32, max_src_w=32)
    eng.load_recognizer(port.build_crnn(seed=0).state_dict())
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in
             [(40, 200), (17, 33), (32, 128), (64, 256), (11, 300), (90, 45), (32, 100), (5, 7)]]
    eng.recognize_crops(crops)
    x = eng.debug_tensor("crops", len(crops))
    for i, c in enumerate(crops):
        want = cv2.resize(c, (128, 32)).transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
        restated = port.cv_resize_linear_restated(c, 32, 128).transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
        assert np.array_equal(x[i], restated), i                      # bit-exact to the portable OpenCV arithmetic
        assert np.abs(x[i] - want).max() <= 1.0 / 255.0 + 1e-7, i     # <= 1 LSB from the installed wheel


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-3), (T16, 5e-3 if T16 == "fp16" else 4e-2)])
def test_crnn_logits(E, port, dtype, tol):
    g = load_golden("crnn")
    net = port.build_crnn(seed=0)
    eng = E.Engine(det_h=32, det_w=32, crop_w=128, max_batch=1, max_boxes=64, max_src_h=32, max_src_w=32, dtype=dtype)
    eng.load_recognizer(net.state_dict())
    out = eng.crnn_forward(g["inputs"])
    ref = g["logits"]
    assert out.shape == ref.shape
    print("crnn %s: max |dlogit| %.2e of max |logit| %.2f" % (dtype, np.abs(out - ref).max(), np.abs(ref).max()))
    assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()
    if dtype == "fp32":
        crops = [g["crop%d" % i] for i in range(int(g["n"]))]
        ids, lens, conf, logits = eng.recognize_crops(crops, want_logits=True)
        res = golden_json(g["results"])
        for i, r in enumerate(res):
            assert E.ids_to_text(ids[i, :lens[i]].tolist()) == r["text"]
            assert conf[i] == pytest.approx(r["confidence"], abs=2e-3)


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), (T16, 1e-3)])
def test_crnn_w100_shapes(E, port, dtype, tol):
    """BASELINE configs[2] words the crops as 32x100 (T=24); measured round 1: 6e-8 (fp32) / 2.6e-4 (bf16)."""
    net = port.build_crnn(seed=3)
    eng = E.Engine(det_h=32, det_w=32, crop_w=100, max_batch=1, max_boxes=64, max_src_h=32, max_src_w=32, dtype=dtype)
    eng.load_recognizer(net.state_dict())
    x = np.random.default_rng(1).random((5, 3, 32, 100)).astype(np.float32)
    out = eng.crnn_forward(x)
    with torch.no_grad():
        ref = net(torch.from_numpy(x)).numpy()
    assert out.shape == ref.shape == (5, 24, 97)
    print("crnn w100 %s: max |dlogit| %.2e" % (dtype, np.abs(out - ref).max()))
    assert np.abs(out - ref).max() <= tol


def test_ctc_golden_bit_exact(E, port):
    g = load_golden("ctc")
    eng = E.Engine(det_h=32, det_w=32, max_batch=1, max_src_h=32, max_src_w=32)
    for i in range(int(g["n"])):
        logits = g["%d_logits" % i]
        want_text = bytes(g["%d_text" % i].tolist()).decode()
        ids, lens, conf = eng.ctc_decode(logits, is_prob=False)
        assert E.ids_to_text(ids[0, :lens[0]].tolist()) == want_text, i
        assert conf[0] == pytest.approx(float(g["%d_conf" % i]), abs=1e-6)
        p = torch.softmax(torch.from_numpy(logits), dim=1).numpy()
        ids2, lens2, conf2 = eng.ctc_decode(p, is_prob=True)
        assert ids2[0, :lens2[0]].tolist() == ids[0, :lens[0]].tolist()
    ids, lens, conf = eng.ctc_decode(g["tie_probs"], is_prob=True)
    assert E.ids_to_text(ids[0, :lens[0]].tolist()) == bytes(g["tie_text"].tolist()).decode()
    assert conf[0] == pytest.approx(float(g["tie_conf"]), abs=1e-7)


def test_ctc_random_vs_oracle_ids(E, port):
    rng = np.random.default_rng(4)
    eng = E.Engine(det_h=32, det_w=32, max_batch=1, max_src_h=32, max_src_w=32)
    for T in (1, 10, 24, 31, 64):
        logits = (rng.standard_normal((64, T, 97)) * 3).astype(np.float32)
        logits[:, :, 0] += 2.0                      # plenty of blanks
        logits[::3, :, 96] += 2.5                   # and <unk>
        ids, lens, conf = eng.ctc_decode(logits, is_prob=False)
        for b in range(64):
            p = torch.softmax(torch.from_numpy(logits[b]), dim=1)
            text, c, want = port.decode_prediction(p)
            assert ids[b, :lens[b]].tolist() == want, (T, b)
            assert conf[b] == pytest.approx(c, abs=1e-6)


# ---------------------------------------------------------------- whole path
def test_full_pipeline_vs_oracle(E, port):
    det = port.build_dbnet("resnet18", seed=0)
    rec = port.build_crnn(seed=0)
    h, w = 256, 1280
    n = 3
    frames = port.synthetic_frames(n, 288, 1440, seed=11)
    bias = port.planted_logit_bias(n, h, w, seed=6, boxes=10)
    eng = E.Engine(backbone=18, det_h=h, det_w=w, max_batch=n, max_boxes=64, max_src_h=288, max_src_w=1440)
    eng.load_detector(det.state_dict())
    eng.load_recognizer(rec.state_dict())
    b = torch.from_numpy(bias).cuda()
    r, c = eng.run_batch(list(frames), thr=0.5, recognize=True, logit_bias_dev=b.data_ptr())
    assert eng.overflow() == 0
    for i in range(n):
        mine = E.records_to_detections(r[i], int(c[i]), True)
        ref = port.process_frame(det, rec, frames[i], 0.5, h, w, 128, torch.from_numpy(bias[i:i + 1])[:, None],
                                 per_crop=False)
        assert len(ref) >= 5
        match_boxes(mine, ref)
        rd = {tuple(d["bbox"]): d for d in ref}
        same_text = 0
        for d in mine:
            q = rd.get(tuple(d["bbox"]))
            if q is None:
                continue
            assert d["confidence"] == pytest.approx(q["detection_confidence"], abs=1e-4)
            same_text += d["ids"] == q["ids"]
        # token ids are bit-exact for identical logits (test_ctc_*); end to end, fp32 reassociation may flip a
        # near-tie argmax on random-init weights, so demand agreement on the bulk, not on every crop
        assert same_text >= 0.8 * len(mine)


def test_golden_pipeline_frame(E, port):
    g = load_golden("pipeline")
    regions = golden_json(g["regions"])
    from video_text_detection_system_b200 import TextDetector, TextRecognizer
    D = TextDetector(backbone="resnet18", pretrained=False, dtype="fp32")
    R = TextRecognizer(use_transformer=False, dtype="fp32")
    R.model.load_state_dict(port.build_crnn(seed=0).state_dict())
    pm = g["planted_map"]
    D.model.forward = lambda x: {"probability": torch.from_numpy(pm)[None, None], "threshold": torch.zeros(1, 1, 640, 640)}
    frame = g["frame"]
    dets = D.detect(frame, 0.5)
    assert sorted(tuple(d["bbox"]) for d in dets) == sorted(tuple(r["bbox"]) for r in regions)
    by = {tuple(r["bbox"]): r for r in regions}
    for d in dets:
        r = by[tuple(d["bbox"])]
        assert d["polygon"] == r["polygon"]
        assert d["confidence"] == pytest.approx(r["detection_confidence"], rel=1e-5)
        x1, y1, x2, y2 = d["bbox"]
        t = R.recognize(frame[y1:y2, x1:x2])
        assert t["text"] == r["text"]
        assert t["confidence"] == pytest.approx(r["recognition_confidence"], abs=2e-3)
