"""Host side of the device overlay: the committed glyph table is what OpenCV renders today, and the draw list is built as
the reference formats its labels (app/services/processing_service.py:198)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_glyph_table_matches_opencv():
    gen = os.path.join(ROOT, "video_text_detection_system_b200", "csrc", "gen_overlay_atlas.py")
    assert subprocess.run([sys.executable, gen, "--check"]).returncode == 0


def test_glyph_table_reproduces_puttext_on_the_host():
    """The blit rule of csrc/overlay.cu, restated in numpy over the same table, against cv2.putText for every printable byte
    at both pen phases."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "video_text_detection_system_b200", "csrc"))
    import gen_overlay_atlas as g
    widths, th, base, cells = g.tables()
    assert (th, base) == (12, 5)
    rng = np.random.default_rng(0)
    for _ in range(40):
        text = "".join(chr(c) for c in rng.integers(32, 127, int(rng.integers(1, 14))))
        x, y = int(rng.integers(3, 30)), int(rng.integers(20, 40))
        want = np.full((64, 320), 255, np.uint8)
        cv2.putText(want, text, (x, y), cv2.FONT_HERSHEY_SIMPLEX, 0.5, 0, 1)
        got = np.full((64, 320), 255, np.uint8)
        pen2 = 2 * x
        for ch in text:
            c = ord(ch) - 32
            for r in range(g.CELL_H):
                bits = int(cells[c, pen2 & 1, r])
                for b in range(g.CELL_W):
                    if bits >> b & 1:
                        got[y + g.ROW0 + r, (pen2 >> 1) + b] = 0
            pen2 += widths[c]
        assert np.array_equal(got, want), text
        (tw, _), _ = cv2.getTextSize(text, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 1)
        assert tw == int(np.rint((pen2 - 2 * x) * 0.5 + 1.0))


def test_clipped_glyph_patches_reproduce_puttext_at_the_frame_border():
    """The rule csrc/overlay.cu applies to a glyph cut by one frame border (plain cell cropped, unless the table holds a cell
    for this glyph / phase / border / distance), restated in numpy, against cv2.putText on labels pushed through each border."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, "video_text_detection_system_b200", "csrc"))
    import gen_overlay_atlas as g
    widths, th, base, cells = g.tables()
    keys, rows = g.clip_patches(widths, cells)
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    patch = dict(zip(keys, rows))

    def draw(h, w, text, x, y):
        im = np.full((h, w), 255, np.uint8)
        pen2 = 2 * x
        for ch in text:
            c, ph, px, ty0 = ord(ch) - 32, pen2 & 1, pen2 >> 1, y + g.ROW0
            kt, kb = max(0, -ty0), max(0, ty0 + g.CELL_H - h)
            kl, kr = max(0, -px), max(0, px + g.CELL_W - w)
            cell = cells[c, ph]
            if (kt > 0) + (kb > 0) + (kl > 0) + (kr > 0) == 1 and kt + kb + kl + kr <= 16:
                side = 0 if kt else 1 if kb else 2 if kl else 3
                cell = patch.get((((c * 2 + ph) * 4 + side) * 17 + kt + kb + kl + kr), cell)
            for r in range(g.CELL_H):
                for b in range(g.CELL_W):
                    if (int(cell[r]) >> b) & 1 and 0 <= ty0 + r < h and 0 <= px + b < w:
                        im[ty0 + r, px + b] = 0
            pen2 += widths[c]
        return im

    rng = np.random.default_rng(0)
    h, w = 64, 200
    for it in range(200):
        text = "".join(chr(c) for c in rng.integers(32, 127, int(rng.integers(1, 10))))
        x, y = [(int(rng.integers(5, 60)), int(rng.integers(-3, 14))), (int(rng.integers(5, 60)), int(rng.integers(h - 6, h + 14))),
                (int(rng.integers(-60, 0)), int(rng.integers(20, 50))), (int(rng.integers(w - 70, w - 2)), int(rng.integers(20, 50)))][it % 4]
        want = np.full((h, w), 255, np.uint8)
        cv2.putText(want, text, (x, y), cv2.FONT_HERSHEY_SIMPLEX, 0.5, 0, 1)
        assert np.array_equal(draw(h, w, text, x, y), want), (text, x, y)


def test_overlay_items_labels():
    from video_text_detection_system_b200._lib import OVERLAY_DTYPE
    from video_text_detection_system_b200.sinks import overlay_items
    items = overlay_items([[{"bbox": [1, 2, 3, 4], "text": "héllo", "detection_confidence": 0.955}, {"bbox": [1, 2, 3]}],
                           [], [{"bbox": [9, 8, 7, 6]}]])
    assert items.dtype == OVERLAY_DTYPE and len(items) == 2
    label = ("%s (%.2f)" % ("héllo", 0.955)).encode("utf-8")
    assert items[0]["frame"] == 0 and items[0]["bbox"].tolist() == [1, 2, 3, 4] and items[0]["label_len"] == len(label)
    assert bytes(items[0]["label"][:len(label)]) == label
    assert items[1]["frame"] == 2 and bytes(items[1]["label"][:items[1]["label_len"]]) == b" (0.00)"
