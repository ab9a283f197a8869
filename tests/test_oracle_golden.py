"""CPU: the oracle (oracle/port.py) against the golden vectors minted from the REFERENCE's own code
(oracle/make_goldens.py).  This is what pins the oracle on machines where /root/reference is absent."""
import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden
from oracle import port


def test_state_dict_checksums_stable():
    # the goldens store no weights: they are the seeded initialisation; detect RNG drift loudly
    from oracle.make_goldens import sd_checksum
    for bb in ("resnet18", "resnet50"):
        g = load_golden("dbnet_" + bb)
        assert sd_checksum(port.build_dbnet(bb, seed=0).state_dict()) == pytest.approx(float(g["sd_checksum"]), rel=1e-9)
    assert sd_checksum(port.build_crnn(seed=0).state_dict()) == pytest.approx(float(load_golden("crnn")["sd_checksum"]), rel=1e-9)


@pytest.mark.parametrize("bb", ["resnet18", "resnet50"])
def test_dbnet_forward_matches_reference_modules(bb):
    g = load_golden("dbnet_" + bb)
    net = port.build_dbnet(bb, seed=0)
    with torch.no_grad():
        r = port.dbnet_forward(net, torch.from_numpy(g["x"]), return_feats=True)
    assert np.abs(r["probability"].numpy() - g["probability"]).max() < 1e-5
    assert np.abs(r["threshold"].numpy() - g["threshold"]).max() < 1e-5
    assert np.abs(r["c5"].numpy() - g["c5"]).max() < 1e-4
    assert np.abs(r["p2"].numpy()[:, ::8, ::2, ::2] - g["p2_s"]).max() < 1e-4


def test_preprocess_and_restated_pillow_resize():
    g = load_golden("preprocess")
    for k in ("structured", "small_random", "hd_gradient"):
        frame = g[k + "_frame"]
        t = port.preprocess(frame, 640, 640)[0].numpy()
        assert np.array_equal(t[:, ::37, ::41], g[k + "_tensor_sample"])
        # the integer restatement the CUDA kernel implements is bit-exact to Pillow
        u8 = port.pillow_resize_restated(np.ascontiguousarray(frame[:, :, ::-1]), 640, 640)
        assert np.array_equal(u8.transpose(2, 0, 1), g[k + "_resized_rgb_u8"])
        assert np.array_equal(port.preprocess_restated(frame, 640, 640), t)


def test_restated_pillow_resize_size_sweep():
    """The integer restatement the CUDA preprocess kernel is held to, against the installed Pillow over a seeded sweep
    of shapes: up- and down-scaling on either axis, 1-pixel sources, odd sizes, the BASELINE frame sizes (scaled down
    4x to keep the pure-NumPy restatement quick).  Plus its stated limit: sources taller than 100x their width."""
    from PIL import Image
    rng = np.random.default_rng(123)
    sizes = [(270, 480, 184, 328), (180, 320, 184, 328), (540, 960, 544, 960), (160, 160, 160, 160), (64, 64, 640, 640),
             (641, 639, 640, 640), (1, 1, 32, 32), (2, 3, 32, 64), (333, 777, 96, 160), (37, 53, 64, 32), (300, 3, 32, 256),
             (3, 400, 32, 64), (100, 10, 50, 40), (64, 16, 32, 32)]
    for _ in range(16):
        sizes.append((int(rng.integers(1, 500)), int(rng.integers(5, 500)), int(rng.integers(1, 12)) * 32,
                      int(rng.integers(1, 12)) * 32))
    for h, w, oh, ow in sizes:
        assert h <= 100 * w
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(port.pillow_resize_restated(img, oh, ow), want), (h, w, oh, ow)
    # beyond the limit Pillow swaps its passes: within 1 LSB, and equal to the restatement run vertical-first
    img = rng.integers(0, 256, (342, 3, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((256, 288), Image.BILINEAR)).astype(int)
    got = port.pillow_resize_restated(img, 288, 256).astype(int)
    assert 0 < np.abs(got - want).max() <= 1
    swapped = port.pillow_resize_restated(port.pillow_resize_restated(img, 288, 3), 288, 256)
    assert np.array_equal(swapped, want)


def test_restated_cv_resize_size_sweep():
    """cv2.resize INTER_LINEAR to 32 x {128,100} (text_recognizer.py:118) against the portable fixed-point restatement the
    crop kernel implements: within 1 LSB of the installed wheel for crops of any shape (the wheel's SIMD path differs
    from OpenCV's own portable path by at most that)."""
    import cv2
    rng = np.random.default_rng(9)
    shapes = [(1, 1), (1, 500), (500, 1), (11, 11), (12, 300), (31, 127), (33, 129), (32, 128), (32, 100), (2, 2), (300, 900)]
    shapes += [(int(rng.integers(1, 400)), int(rng.integers(1, 900))) for _ in range(30)]
    for h, w in shapes:
        c = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for cw in (128, 100):
            a = port.cv_resize_linear_restated(c, 32, cw).astype(int)
            b = cv2.resize(c, (cw, 32)).astype(int)
            assert np.abs(a - b).max() <= 1, (h, w, cw)


def _cases():
    g = load_golden("postprocess")
    return sorted({k[:-4] for k in g.files if k.endswith("_map")})


@pytest.mark.parametrize("case", _cases())
def test_post_process_matches_reference(case):
    g = load_golden("postprocess")
    ow, oh, thr = g[case + "_args"]
    mine = port.post_process(g[case + "_map"], int(ow), int(oh), float(thr))
    ref = golden_json(g[case + "_dets"])
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a["bbox"] == b["bbox"] and a["polygon"] == b["polygon"]
        assert a["confidence"] == pytest.approx(b["confidence"], rel=1e-6, nan_ok=True)


def test_known_semantics():
    g = load_golden("postprocess")
    # ring + island: the island is nested in a hole -> not RETR_EXTERNAL
    assert len(golden_json(g["ring_island_dets"])) == 1
    # 11x11 block: contour area 100 passes `< 100`, but its 10-px box fails `> 10`; 10x11 (area 90) is dropped
    # by the area filter; only the 30x60 block survives
    assert len(golden_json(g["area_filter_dets"])) == 1
    # strict '>': a plane of exactly 0.5 is background
    assert len(golden_json(g["strict_gt_dets"])) == 1
    assert len(golden_json(g["diag_touch_dets"])) == 1


def test_decode_matches_reference():
    g = load_golden("ctc")
    for i in range(int(g["n"])):
        p = torch.softmax(torch.from_numpy(g["%d_logits" % i]), dim=1)
        text, conf, ids = port.decode_prediction(p)
        assert text == bytes(g["%d_text" % i].tolist()).decode()
        assert conf == pytest.approx(float(g["%d_conf" % i]), abs=1e-7)
        assert "".join(port.CHARS[i - 1] for i in ids) == text
    text, conf, _ = port.decode_prediction(torch.from_numpy(g["tie_probs"]))
    assert text == bytes(g["tie_text"].tolist()).decode() and conf == pytest.approx(float(g["tie_conf"]))


def test_decode_reference_quirks():
    def run(seq):
        p = torch.full((len(seq), 97), 1e-3)
        for t, s in enumerate(seq):
            p[t, s] = 0.9
        return port.decode_prediction(p)[0]
    a, b = 11, 12          # 'a', 'b'
    assert run([a, 0, a, b]) == "ab"           # blank does not reset prev (canonical CTC gives "aab")
    assert run([a, a, b, b, 0]) == "ab"
    assert run([96, a, 96, a]) == "aa"         # <unk> dropped but becomes prev
    assert run([a, b, a]) == "aba"
    assert port.decode_prediction(torch.full((4, 97), 0.0).index_fill_(1, torch.tensor([0]), 1.0))[:2] == ("", 0.0)


def test_crnn_matches_reference():
    g = load_golden("crnn")
    net = port.build_crnn(seed=0)
    crops = [g["crop%d" % i] for i in range(int(g["n"]))]
    x = port.crnn_inputs(crops)
    assert np.array_equal(x.numpy(), g["inputs"])
    with torch.no_grad():
        logits = net(x).numpy()
    assert np.abs(logits - g["logits"]).max() < 1e-5
    res = port.recognize_batch(net, crops)
    for a, b in zip(res, golden_json(g["results"])):
        assert a["text"] == b["text"] and a["confidence"] == pytest.approx(b["confidence"], abs=1e-6)


def test_cv_resize_restatement_within_one_lsb():
    import cv2
    rng = np.random.default_rng(0)
    for h, w in [(40, 200), (17, 33), (32, 128), (64, 256), (90, 45)]:
        c = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        a = port.cv_resize_linear_restated(c, 32, 128).astype(int)
        b = cv2.resize(c, (128, 32)).astype(int)
        assert np.abs(a - b).max() <= 1


def test_pipeline_frame_matches_reference():
    g = load_golden("pipeline")
    regions = golden_json(g["regions"])
    frame, pm = g["frame"], g["planted_map"]
    dets = port.post_process(pm, frame.shape[1], frame.shape[0], 0.5)
    assert [d["bbox"] for d in dets] == [r["bbox"] for r in regions]
    rec = port.build_crnn(seed=0)
    for d, r in zip(dets, regions):
        x1, y1, x2, y2 = d["bbox"]
        t = port.recognize_batch(rec, [frame[y1:y2, x1:x2]])[0]
        assert t["text"] == r["text"]
        assert t["confidence"] == pytest.approx(r["recognition_confidence"], abs=1e-6)
        assert d["confidence"] == pytest.approx(r["detection_confidence"], rel=1e-6)


def test_planted_plane_gives_about_fifty_boxes():
    b = port.planted_logit_bias(1, 736, 1312, seed=0, boxes=50)[0]
    pm = 1.0 / (1.0 + np.exp(-b))
    dets = port.post_process(pm.astype(np.float32), 1920, 1080, 0.5, 736, 1312)
    assert 45 <= len(dets) <= 50
