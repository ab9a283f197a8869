"""CPU check of the product's contour / hull / min-area-rect code (csrc/box_geom.cuh) against cv2 itself.
The same header is compiled into the CUDA box-extraction kernels (boxes.cu)."""
import ctypes

import cv2
import numpy as np
import pytest

from conftest import load_golden


class HhComp(ctypes.Structure):
    _fields_ = [("start", ctypes.c_int32), ("external", ctypes.c_int32), ("area2", ctypes.c_int64),
                ("rect", ctypes.c_float * 5), ("box", ctypes.c_float * 8), ("nhull", ctypes.c_int32),
                ("pad", ctypes.c_int32)]


def run_harness(hh, mask):
    h, w = mask.shape
    cap = 1 << 18
    buf = (HhComp * cap)()
    hh.hh_components.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    m = np.ascontiguousarray(mask.astype(np.uint8))
    n = hh.hh_components(m.ctypes.data, h, w, buf, cap)
    assert n >= 0
    return [buf[i] for i in range(n)]


def cv_externals(mask):
    cs, _ = cv2.findContours((mask > 0).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cn, _ = cv2.findContours((mask > 0).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    out = {}
    for c, cfull in zip(cs, cn):
        pts = cfull[:, 0, :]
        i = np.lexsort((pts[:, 0], pts[:, 1]))[0]           # raster-first border pixel
        start = int(pts[i, 1]) * mask.shape[1] + int(pts[i, 0])
        out[start] = c
    return out


def masks():
    rng = np.random.default_rng(11)
    g = load_golden("postprocess")
    names = sorted({k[:-4] for k in g.files if k.endswith("_map")})
    for nme in names:
        thr = float(g[nme + "_args"][2])
        yield nme, (g[nme + "_map"] > thr)
    for i in range(6):
        f = cv2.GaussianBlur(rng.random((200, 300)).astype(np.float32), (0, 0), 2 + i)
        yield "blur%d" % i, f > np.quantile(f, 0.5 + 0.05 * i)
    yield "noise", rng.random((96, 128)) > 0.55
    yield "full", np.ones((40, 50), bool)
    yield "empty", np.zeros((40, 50), bool)
    m = np.zeros((64, 64), bool); m[10:50, 10:50] = True; m[20:40, 20:40] = False; m[25:35, 25:35] = True
    m[28:32, 28:32] = False; m[29:31, 29:31] = True
    yield "nested", m


@pytest.mark.parametrize("name,mask", list(masks()), ids=[n for n, _ in masks()])
def test_components_match_cv2(host_harness, name, mask):
    comps = run_harness(host_harness, mask)
    ext = {c.start: c for c in comps if c.external}
    ref = cv_externals(mask)
    assert set(ext) == set(ref), "RETR_EXTERNAL membership differs"
    for start, contour in ref.items():
        c = ext[start]
        assert abs(c.area2) / 2.0 == cv2.contourArea(contour)
        if cv2.contourArea(contour) < 100:
            continue
        (cx, cy), (rw, rh), ang = cv2.minAreaRect(contour)
        box = cv2.boxPoints(((cx, cy), (rw, rh), ang))
        mine = np.array(list(c.box), np.float32).reshape(4, 2)
        # bit-exact: same hull order, same calipers decisions, same float32 arithmetic as cv2
        assert np.array_equal(mine, box), (name, start, mine, box)


def shape_masks():
    """Shapes the blur/noise masks above do not produce: rotated rectangles (the production shape) at regular and
    degenerate angles, squares and diamonds (calipers ties), ellipses, dilated text, thin and 1-pixel diagonal lines,
    blobs touching the image border, the 10x11 / 11x11 area-threshold pair."""
    rng = np.random.default_rng(42)
    for ang in list(range(-90, 91, 5)) + [45.0, -45.0, 26.565, 63.435, 1e-3, 89.999]:
        for bw, bh in ((80, 24), (40, 40), (101, 13)):
            m = np.zeros((160, 200), np.uint8)
            cv2.fillPoly(m, [np.round(cv2.boxPoints(((100, 80), (bw, bh), float(ang)))).astype(np.int32)], 1)
            yield "rect a=%s %dx%d" % (ang, bw, bh), m > 0
    for k in range(20):
        m = np.zeros((300, 400), np.uint8)
        for _ in range(12):
            box = cv2.boxPoints(((rng.uniform(30, 370), rng.uniform(30, 270)),
                                 (rng.uniform(15, 100), rng.uniform(11, 40)), rng.uniform(-90, 90)))
            cv2.fillPoly(m, [np.round(box).astype(np.int32)], 1)
        yield "multi%d" % k, m > 0
    for k in range(12):
        m = np.zeros((200, 300), np.uint8)
        cv2.ellipse(m, (150, 100), (int(rng.integers(8, 120)), int(rng.integers(8, 80))), float(rng.uniform(0, 180)),
                    0, 360, 1, -1)
        yield "ellipse%d" % k, m > 0
    for k in range(12):
        m = np.zeros((120, 500), np.uint8)
        cv2.putText(m, "Text %d gq!" % k, (10, 80), cv2.FONT_HERSHEY_SIMPLEX, rng.uniform(1, 3), 1, int(rng.integers(2, 8)))
        yield "text%d" % k, (cv2.dilate(m, np.ones((5, 9), np.uint8)) if k % 2 else m) > 0
    for k in range(8):
        m = np.zeros((200, 300), np.uint8)
        cv2.line(m, (int(rng.integers(0, 300)), int(rng.integers(0, 200))),
                 (int(rng.integers(0, 300)), int(rng.integers(0, 200))), 1, int(rng.integers(1, 4)))
        yield "line%d" % k, m > 0
    for k in range(8):
        m = np.zeros((100, 150), np.uint8)
        m[:int(rng.integers(5, 40)), :int(rng.integers(20, 150))] = 1
        m[-int(rng.integers(5, 40)):, -int(rng.integers(20, 150)):] = 1
        yield "border%d" % k, m > 0
    for hh_, name in ((10, "10x11"), (11, "11x11")):
        m = np.zeros((64, 64), bool)
        m[10:21, 10:10 + hh_] = True
        yield name, m
    m = np.zeros((64, 64), bool)
    m[np.arange(10, 50), np.arange(10, 50)] = True
    yield "diagonal-1px", m
    yy, xx = np.mgrid[0:101, 0:101]
    yield "diamond", (np.abs(xx - 50) + np.abs(yy - 50)) <= 30


def test_shape_sweep_matches_cv2(host_harness):
    boxes = 0
    for name, mask in shape_masks():
        comps = run_harness(host_harness, mask)
        ext = {c.start: c for c in comps if c.external}
        ref = cv_externals(mask)
        assert set(ext) == set(ref), name
        for start, contour in ref.items():
            c = ext[start]
            assert abs(c.area2) / 2.0 == cv2.contourArea(contour), (name, start)
            if cv2.contourArea(contour) < 100:
                continue
            box = cv2.boxPoints(cv2.minAreaRect(contour))
            assert np.array_equal(np.array(list(c.box), np.float32).reshape(4, 2), box), (name, start)
            boxes += 1
    assert boxes >= 300


def test_resize_coefficient_tables_match_the_oracle(host_harness):
    """csrc/resize_tab.h (the tables vtd_preprocess uploads for the Pillow-exact resize kernel) against the oracle's
    pillow_coeffs -- which the size sweeps of tests/test_oracle_golden.py hold to Pillow itself -- for the BASELINE frame and
    detector sizes and a seeded sweep of up- and down-scaling ratios: first tap, tap count and every 22-bit weight."""
    from oracle import port
    rng = np.random.default_rng(3)
    pairs = [(1080, 736), (1920, 1312), (2160, 2176), (3840, 3840), (720, 736), (1280, 1312), (480, 640), (640, 640),
             (360, 320), (540, 480), (1, 32), (2, 64), (3, 256), (5000, 32), (4320, 736), (7680, 1312), (641, 640), (639, 640)]
    pairs += [(int(rng.integers(1, 4000)), int(rng.integers(1, 70)) * 32) for _ in range(60)]
    host_harness.hh_resize_tab.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_void_p]
    for n_in, n_out in pairs:
        lo_w, cnt_w, kk_w = port.pillow_coeffs(n_in, n_out)
        ks = kk_w.shape[1]
        lo = np.zeros(n_out, np.int32)
        cnt = np.zeros(n_out, np.int32)
        kk = np.zeros((n_out, ks), np.int32)
        mx = ctypes.c_int(0)
        got_ks = host_harness.hh_resize_tab(n_in, n_out, lo.ctypes.data, cnt.ctypes.data, kk.ctypes.data, ks, ctypes.byref(mx))
        assert got_ks == ks, (n_in, n_out)
        assert np.array_equal(lo, lo_w) and np.array_equal(cnt, cnt_w), (n_in, n_out)
        assert np.array_equal(kk, kk_w), (n_in, n_out, int(np.abs(kk - kk_w).max()))
        assert mx.value == int(cnt_w.max())
        assert np.all(np.abs(kk.sum(1) - (1 << 22)) <= ks)            # rows sum to 1.0 in 22-bit fixed point, up to rounding


def test_unclip_matches_the_stated_formula(host_harness):
    """north_star's unclip (an extension; ratio 1.0 = the reference): the product's unclip_rect + box_points against the
    oracle's statement of the same formula (d = w*h*ratio / (2(w+h)), float32) followed by cv2.boxPoints, bit for bit,
    over random rotated rects; ratio <= 1 leaves the rect untouched."""
    import cv2
    from oracle import port
    host_harness.hh_unclip_box.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    host_harness.hh_unclip_box.restype = None
    rng = np.random.default_rng(5)
    for i in range(400):
        rect = ((float(np.float32(rng.uniform(20, 1200))), float(np.float32(rng.uniform(20, 700)))),
                (float(np.float32(rng.uniform(5, 300))), float(np.float32(rng.uniform(5, 120)))),
                float(np.float32(rng.uniform(-90, 0))))
        for ratio in (1.0, 0.5, 1.5, 2.0):
            r5 = np.asarray([rect[0][0], rect[0][1], rect[1][0], rect[1][1], rect[2]], np.float32)
            o5, o8 = np.zeros(5, np.float32), np.zeros(8, np.float32)
            host_harness.hh_unclip_box(r5.ctypes.data, ratio, o5.ctypes.data, o8.ctypes.data)
            want = port.unclip_rect(rect, ratio)
            assert (o5[2], o5[3]) == (np.float32(want[1][0]), np.float32(want[1][1])), (rect, ratio)
            if ratio <= 1.0:
                assert np.array_equal(o5, r5)
            else:
                assert o5[2] > r5[2] and o5[3] > r5[3]
            assert np.array_equal(o8.reshape(4, 2), cv2.boxPoints(want)), (rect, ratio)
