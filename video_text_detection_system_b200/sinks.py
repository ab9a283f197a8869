"""Result sinks: the step right AFTER the hot path (host half).

The reference turns the pipeline's result dictionary into
  * CSV                (app/services/processing_service.py:59-88),
  * XML                (app/services/processing_service.py:90-137),
  * database rows      (app/tasks/video_processing.py:169-216: one Pydantic object per frame and per detection),
  * an annotated frame (app/services/processing_service.py:188-218).
At the rates the device path delivers (16 frames x ~50 detections every ~5 ms) an object per detection is the
bottleneck, so the text formats are written as strings in one pass per frame, and the database rows are plain dicts in
the bulk-insert shape.  Outputs are character-for-character those of the reference's functions for the same result
dictionary (tests/test_sinks.py checks them against goldens minted from the reference's own code, and against
xml.etree / csv on hostile strings).  The annotated frame exists twice: `draw_detections` (OpenCV on the host, the reference's
own calls) and `OverlayRenderer` / `draw_detections_batch` (csrc/overlay.cu through vtd_draw_detections: a batch of frames
drawn in HBM, pixel-identical to the host version for labels inside the frame).
"""
from __future__ import annotations

import csv
import io
import logging
from typing import Any, Dict, Iterable, List, Tuple

import numpy as np

logger = logging.getLogger(__name__)

CSV_HEADER = ("frame_number", "timestamp", "text", "bbox_x1", "bbox_y1", "bbox_x2", "bbox_y2",
              "detection_confidence", "recognition_confidence")
MODEL_NAME, MODEL_VERSION = "DBNet-CRNN", "1.0.0"              # video_processing.py:202-203


# ------------------------------------------------------------------------------------------------ CSV
def _csv_rows(results_data: Dict[str, Any]) -> Iterable[Tuple]:
    for fr in results_data.get("results", []):
        number, stamp = fr.get("frame_number", 0), fr.get("timestamp", 0.0)
        for d in fr.get("detections", []):
            x1, y1, x2, y2 = d.get("bbox", [0, 0, 0, 0])[:4]
            yield (number, stamp, d.get("text", ""), x1, y1, x2, y2,
                   d.get("detection_confidence", 0.0), d.get("recognition_confidence", 0.0))


def export_results_csv(results_data: Dict[str, Any]) -> str:
    """processing_service.py:59-88.  '' on error, as the reference."""
    try:
        out = io.StringIO()
        w = csv.writer(out)
        w.writerow(CSV_HEADER)
        w.writerows(_csv_rows(results_data))
        return out.getvalue()
    except Exception as e:
        logger.error(f"CSV export failed: {e}")
        return ""


# ------------------------------------------------------------------------------------------------ XML
_ATTR_ESC = {"&": "&amp;", "<": "&lt;", ">": "&gt;", '"': "&quot;", "\r": "&#13;", "\n": "&#10;", "\t": "&#09;"}
_TEXT_ESC = {"&": "&amp;", "<": "&lt;", ">": "&gt;"}
_ATTR_TABLE = str.maketrans(_ATTR_ESC)
_TEXT_TABLE = str.maketrans(_TEXT_ESC)


def _attr(v: Any) -> str:
    """xml.etree.ElementTree's attribute escaping (ElementTree._escape_attrib)."""
    return str(v).translate(_ATTR_TABLE)


def export_results_xml(results_data: Dict[str, Any]) -> str:
    """processing_service.py:90-137: <video_text_detection><summary>..</summary><frames><frame number timestamp>
    <object transcription detection_confidence recognition_confidence><Point x y/> x4 (bbox corners clockwise from
    top-left)</object>..  Serialised exactly as ET.tostring(root, encoding='unicode') does (empty elements as
    '<tag />', ElementTree escaping).  '' on error."""
    try:
        parts: List[str] = ["<video_text_detection>"]
        summary = results_data.get("summary", {})
        if summary:
            parts.append("<summary>")
            for key, value in summary.items():
                text = str(value)
                parts.append("<%s>%s</%s>" % (key, text.translate(_TEXT_TABLE), key) if text else "<%s />" % key)
            parts.append("</summary>")
        else:
            parts.append("<summary />")
        frames = results_data.get("results", [])
        parts.append("<frames>" if frames else "<frames />")
        for fr in frames:
            head = '<frame number="%s" timestamp="%s"' % (_attr(fr.get("frame_number", 0)),
                                                           _attr(fr.get("timestamp", 0.0)))
            dets = fr.get("detections", [])
            if not dets:
                parts.append(head + " />")
                continue
            parts.append(head + ">")
            for d in dets:
                b = d.get("bbox", [0, 0, 0, 0])
                x1, y1, x2, y2 = _attr(b[0]), _attr(b[1]), _attr(b[2]), _attr(b[3])
                parts.append('<object transcription="%s" detection_confidence="%s" recognition_confidence="%s">'
                             '<Point x="%s" y="%s" /><Point x="%s" y="%s" /><Point x="%s" y="%s" />'
                             '<Point x="%s" y="%s" /></object>'
                             % (_attr(d.get("text", "")), _attr(d.get("detection_confidence", 0.0)),
                                _attr(d.get("recognition_confidence", 0.0)), x1, y1, x2, y1, x2, y2, x1, y2))
            parts.append("</frame>")
        if frames:
            parts.append("</frames>")
        parts.append("</video_text_detection>")
        return "".join(parts)
    except Exception as e:
        logger.error(f"XML export failed: {e}")
        return ""


# ------------------------------------------------------------------------------------------------ database rows
def database_rows(video_id: int, results: Dict[str, Any]) -> Tuple[List[Dict[str, Any]], List[Dict[str, Any]]]:
    """video_processing.py:169-216 without the per-row Pydantic objects: (frame rows, detection rows) as plain dicts in
    the FrameCreate / TextDetectionCreate field layout (app/database/schemas.py:75-107).  A detection row carries
    'frame_number' in place of 'frame_id': the caller maps it after the bulk insert of the frames returns their ids
    (frame_mapping, :187-188).  Raises on malformed input, as the reference does (:208-210)."""
    info = results["video_info"]
    width, height = info.get("width", 640), info.get("height", 480)
    frames, dets = [], []
    for fr in results["results"]:
        number = fr["frame_number"]
        frames.append({"video_id": video_id, "frame_number": number, "timestamp": fr["timestamp"],
                       "file_path": f"frame_{number:04d}.jpg", "width": width, "height": height})
        for d in fr["detections"]:
            b = d["bbox"]
            dets.append({"frame_number": number, "text_content": d["text"], "confidence": d["detection_confidence"],
                         "bbox_x1": b[0], "bbox_y1": b[1], "bbox_x2": b[2], "bbox_y2": b[3],
                         "model_name": MODEL_NAME, "model_version": MODEL_VERSION})
    return frames, dets


# ------------------------------------------------------------------------------------------------ overlay (host)
def draw_detections(frame: np.ndarray, detections: List[Dict[str, Any]]) -> np.ndarray:
    """processing_service.py:188-218: green 2-px box, filled label plate above it, black label 'text (0.xx)'.
    Draws in place and returns the frame, as the reference."""
    import cv2
    green, black, font = (0, 255, 0), (0, 0, 0), cv2.FONT_HERSHEY_SIMPLEX
    for d in detections:
        box = d.get("bbox", [])
        if len(box) != 4:
            continue
        x1, y1, x2, y2 = box
        label = "%s (%.2f)" % (d.get("text", ""), d.get("detection_confidence", 0.0))
        (tw, th), _ = cv2.getTextSize(label, font, 0.5, 1)
        cv2.rectangle(frame, (x1, y1), (x2, y2), green, 2)
        cv2.rectangle(frame, (x1, y1 - th - 10), (x1 + tw, y1), green, -1)
        cv2.putText(frame, label, (x1, y1 - 5), font, 0.5, black, 1)
    return frame


# ---------------------------------------------------------------------------------------------- overlay (device)
def overlay_items(detections_per_frame: List[List[Dict[str, Any]]]) -> np.ndarray:
    """The draw list of vtd_draw_detections for a batch: one record per detection with a 4-element bbox, label
    "%s (%.2f)" % (text, detection_confidence) as processing_service.py:198 formats it, UTF-8 bytes (OpenCV draws '?' for
    every byte outside 32..126)."""
    from ._lib import OVERLAY_DTYPE, OVERLAY_LABEL_MAX
    rows = []
    for f, dets in enumerate(detections_per_frame):
        for d in dets:
            box = d.get("bbox", [])
            if len(box) != 4:
                continue
            label = ("%s (%.2f)" % (d.get("text", ""), d.get("detection_confidence", 0.0))).encode("utf-8")
            if len(label) > OVERLAY_LABEL_MAX:
                raise ValueError("label of %d bytes exceeds the overlay's %d" % (len(label), OVERLAY_LABEL_MAX))
            rows.append((f, [int(v) for v in box], label))
    items = np.zeros(len(rows), OVERLAY_DTYPE)
    for i, (f, box, label) in enumerate(rows):
        items[i]["frame"] = f
        items[i]["bbox"] = box
        items[i]["label_len"] = len(label)
        items[i]["label"][:len(label)] = np.frombuffer(label, np.uint8)
    return items


class OverlayRenderer:
    """Draws the detections of a batch of equally sized BGR frames on the device, in place (vtd_draw_detections).  Owns a
    small context sized for the frames it is given; pass `engine=` to draw with an existing one (its max_batch / max_src
    must cover the frames)."""

    def __init__(self, engine=None, device: int = 0, max_batch: int = 16):
        self._engine = engine
        self._own = None
        self._device = int(device)
        self._max_batch = int(max_batch)

    def _eng(self, n: int, h: int, w: int):
        if self._engine is not None:
            return self._engine
        key = (max(n, self._max_batch), h, w)
        if self._own is None or self._own[0][0] < key[0] or self._own[0][1:] != key[1:]:
            from ._lib import Engine
            self._own = (key, Engine(device=self._device, dtype="fp16", det_h=32, det_w=32, max_batch=key[0], max_boxes=64,
                                     max_src_h=h, max_src_w=w))
        return self._own[1]

    def draw(self, frames: List[np.ndarray], detections_per_frame: List[List[Dict[str, Any]]]) -> List[np.ndarray]:
        if len(frames) != len(detections_per_frame):
            raise ValueError("one detection list per frame")
        if not frames:
            return frames
        items = overlay_items(detections_per_frame)
        h, w = frames[0].shape[:2]
        eng = self._eng(len(frames), h, w)
        step = eng.cfg.max_batch
        for first in range(0, len(frames), step):
            part = frames[first:first + step]
            sel = items[(items["frame"] >= first) & (items["frame"] < first + len(part))].copy()
            sel["frame"] -= first
            if len(sel):
                eng.draw_detections(part, sel)
        return frames


_renderer = None


def draw_detections_batch(frames: List[np.ndarray], detections_per_frame: List[List[Dict[str, Any]]]) -> List[np.ndarray]:
    """Device counterpart of [draw_detections(f, d) for f, d in zip(frames, detections_per_frame)]; draws in place."""
    global _renderer
    if _renderer is None:
        _renderer = OverlayRenderer()
    return _renderer.draw(frames, detections_per_frame)


class ResultSinks:
    """The reference's method names (ProcessingService, processing_service.py:59-218) over the functions above."""

    async def export_results_csv(self, results_data: Dict[str, Any]) -> str:
        return export_results_csv(results_data)

    async def export_results_xml(self, results_data: Dict[str, Any]) -> str:
        return export_results_xml(results_data)

    def _draw_detections(self, frame: np.ndarray, detections: List[Dict[str, Any]]) -> np.ndarray:
        return draw_detections(frame, detections)
