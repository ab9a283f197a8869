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
