"""Generates csrc/overlay_atlas.h: the glyph cells the device overlay (csrc/overlay.cu) blits for the label text.

The reference's sink draws `cv2.putText(frame, label, (x1, y1 - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)`
(app/services/processing_service.py:208-216).  OpenCV rasterises a glyph from its pen position in 16.16 fixed point; at scale 0.5
the pen advances by half a Hershey unit per unit of glyph width, so a glyph lands on one of two sub-pixel phases (pen at x.0 or
x.5) and is otherwise translation invariant.  The table therefore holds, for each byte 32..126 and each phase, the 16x17 cell
of pixels OpenCV sets relative to (floor(pen), baseline - 12) -- rendered here by OpenCV itself, so the device reproduces the
reference's pixels without re-implementing its line rasteriser -- plus the glyph advance widths (Hershey units, scale 1) and
the constant text height getTextSize() returns at this scale.

    python video_text_detection_system_b200/csrc/gen_overlay_atlas.py            # rewrites overlay_atlas.h
    python video_text_detection_system_b200/csrc/gen_overlay_atlas.py --check    # exit 1 if the committed header differs
"""
from __future__ import annotations

import os
import sys

import numpy as np

FIRST, LAST = 32, 126
CELL_W, CELL_H, ROW0 = 16, 17, -12            # cell columns pen+0..pen+15, rows baseline-12..baseline+4


def tables():
    import cv2
    font = cv2.FONT_HERSHEY_SIMPLEX
    widths = [cv2.getTextSize(chr(c), font, 1.0, 1)[0][0] - 1 for c in range(FIRST, LAST + 1)]
    (_, th), base = cv2.getTextSize("Ag", font, 0.5, 1)
    ox, oy = 12, 28

    def render(s):
        im = np.full((48, 256), 255, np.uint8)
        cv2.putText(im, s, (ox, oy), font, 0.5, 0, 1)
        return im == 0

    # a glyph of odd width followed by a space moves the pen by k + 1/2 pixels; the space keeps its strokes clear of the next glyph
    odd = next(c for c in range(FIRST, LAST + 1) if widths[c - FIRST] % 2)
    prefix = chr(odd) + " "
    shift2 = widths[odd - FIRST] + widths[0]
    assert shift2 % 2 == 1
    pre = render(prefix)
    cells = np.zeros((LAST - FIRST + 1, 2, CELL_H), np.uint16)
    for c in range(FIRST, LAST + 1):
        m0 = render(chr(c))
        both = render(prefix + chr(c))
        assert not (pre & ~both).any()
        m1 = both & ~pre
        for ph, m, x0 in ((0, m0, ox), (1, m1, ox + shift2 // 2)):
            ys, xs = np.nonzero(m)
            if len(xs):
                assert xs.min() >= x0 and xs.max() < x0 + CELL_W and ys.min() >= oy + ROW0 and ys.max() < oy + ROW0 + CELL_H, chr(c)
            for r in range(CELL_H):
                bits = 0
                for b in range(CELL_W):
                    if m[oy + ROW0 + r, x0 + b]:
                        bits |= 1 << b
                cells[c - FIRST, ph, r] = bits
    return widths, int(th), int(base), cells


SIDES = ("top", "bottom", "left", "right")


def _bits(mask, x0, y0):
    """CELL_H rows of CELL_W bits of a boolean image, cell origin (x0, y0); pixels outside the image read 0."""
    h, w = mask.shape
    out = np.zeros(CELL_H, np.uint16)
    weights = (1 << np.arange(CELL_W)).astype(np.uint32)
    for r in range(CELL_H):
        y = y0 + r
        if 0 <= y < h:
            xs = np.arange(x0, x0 + CELL_W)
            ok = (xs >= 0) & (xs < w)
            row = np.zeros(CELL_W, bool)
            row[ok] = mask[y, xs[ok]]
            out[r] = int((row * weights).sum())
    return out


def clip_patches(widths, cells):
    """Where a glyph stroke crosses the frame border OpenCV clips the segment (clipLine, 16.16 fixed point) before it rasterises
    it, so the pixels that remain can differ from the unclipped glyph's.  For every glyph, phase, border and number k of cell
    rows / columns beyond that border (1..16) the clipped glyph is rendered by OpenCV and kept when it differs from the cropped
    cell: key = (((c - 32) * 2 + phase) * 4 + side) * 17 + k.  The result does not depend on where along the border the glyph
    sits nor on the frame size (checked here on a second frame size)."""
    import cv2
    font = cv2.FONT_HERSHEY_SIMPLEX
    odd = next(c for c in range(FIRST, LAST + 1) if widths[c - FIRST] % 2)
    prefix = chr(odd) + " "
    shift2 = widths[odd - FIRST] + widths[0]

    def clipped(c, ph, h, w, penx, oy):
        im = np.full((h, w), 255, np.uint8)
        if ph == 0:
            cv2.putText(im, chr(c), (penx, oy), font, 0.5, 0, 1)
            return _bits(im == 0, penx, oy + ROW0)
        ox = penx - shift2 // 2
        cv2.putText(im, prefix + chr(c), (ox, oy), font, 0.5, 0, 1)
        pre = np.full((h, w), 255, np.uint8)
        cv2.putText(pre, prefix, (ox, oy), font, 0.5, 0, 1)
        return _bits((im == 0) & ~(pre == 0), penx, oy + ROW0)

    def cropped(c, ph, h, w, penx, oy):
        out = np.zeros(CELL_H, np.uint16)
        for r in range(CELL_H):
            y = oy + ROW0 + r
            if 0 <= y < h:
                keep = 0
                for b in range(CELL_W):
                    if 0 <= penx + b < w and (int(cells[c - FIRST, ph, r]) >> b) & 1:
                        keep |= 1 << b
                out[r] = keep
        return out

    keys, rows = [], []
    for c in range(FIRST, LAST + 1):
        for ph in range(2):
            for side in range(4):
                for k in range(1, CELL_W + 1 if side >= 2 else CELL_H):
                    got = []
                    for (h, w, along) in ((80, 120, 40), (57, 203, 23)):
                        px, oy = ((along, -ROW0 - k), (along, h - 1 - (CELL_H - 1 + ROW0) + k), (-k, along), (w - CELL_W + k, along))[side]
                        got.append((clipped(c, ph, h, w, px, oy), cropped(c, ph, h, w, px, oy)))
                    assert np.array_equal(got[0][0], got[1][0]), (chr(c), ph, SIDES[side], k)
                    if not np.array_equal(got[0][0], got[0][1]):
                        keys.append((((c - FIRST) * 2 + ph) * 4 + side) * 17 + k)
                        rows.append(got[0][0])
    return keys, rows


def header_text() -> str:
    widths, th, base, cells = tables()
    pkeys, prows = clip_patches(widths, cells)
    out = ["// GENERATED by gen_overlay_atlas.py from OpenCV's own rendering of FONT_HERSHEY_SIMPLEX at scale 0.5, thickness 1 -- do not edit.",
           "// cells[c - 32][phase][row]: bit b = pixel (floor(pen) + b, baseline - 12 + row); phase = pen at x.0 / x.5.",
           "#pragma once", "#include <cstdint>", "namespace vtd {",
           "constexpr int OV_FIRST = %d, OV_LAST = %d, OV_CELL_W = %d, OV_CELL_H = %d, OV_ROW0 = %d;" % (FIRST, LAST, CELL_W, CELL_H, ROW0),
           "constexpr int OV_TEXT_H = %d, OV_BASELINE = %d;   // cv2.getTextSize(label, FONT_HERSHEY_SIMPLEX, 0.5, 1)" % (th, base),
           "static const uint8_t OV_WIDTHS[%d] = {%s};" % (len(widths), ", ".join(str(w) for w in widths)),
           "static const uint16_t OV_CELLS[%d][2][%d] = {" % (cells.shape[0], CELL_H)]
    for i in range(cells.shape[0]):
        rows = ["{%s}" % ", ".join("0x%04x" % v for v in cells[i, ph]) for ph in range(2)]
        ch = chr(FIRST + i)
        out.append("  {%s,\n   %s},  // %s" % (rows[0], rows[1], "backslash" if ch == "\\" else repr(ch)))
    out += ["};",
            "// glyphs clipped by ONE frame border whose remaining pixels differ from the cropped cell (see gen_overlay_atlas.py clip_patches):",
            "// OV_PATCH_KEYS sorted; key = (((c - 32) * 2 + phase) * 4 + side) * 17 + k, side 0 top 1 bottom 2 left 3 right, k = cell rows / columns beyond it",
            "constexpr int OV_PATCHES = %d;" % len(pkeys),
            "static const uint32_t OV_PATCH_KEYS[%d] = {%s};" % (len(pkeys), ", ".join(str(k) for k in pkeys)),
            "static const uint16_t OV_PATCH_CELLS[%d][%d] = {" % (len(pkeys), CELL_H)]
    for r in prows:
        out.append("  {%s}," % ", ".join("0x%04x" % int(v) for v in r))
    out += ["};", "}  // namespace vtd", ""]
    return "\n".join(out)


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay_atlas.h")
    text = header_text()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(path) and open(path).read() == text else 1)
    with open(path, "w") as f:
        f.write(text)
    print("wrote", path, len(text), "bytes")
