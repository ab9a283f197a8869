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
xml.etree / csv on hostile strings).  Nothing here touches the device; the overlay drawn on the GPU is not built yet.
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


class ResultSinks:
    """The reference's method names (ProcessingService, processing_service.py:59-218) over the functions above."""

    async def export_results_csv(self, results_data: Dict[str, Any]) -> str:
        return export_results_csv(results_data)

    async def export_results_xml(self, results_data: Dict[str, Any]) -> str:
        return export_results_xml(results_data)

    def _draw_detections(self, frame: np.ndarray, detections: List[Dict[str, Any]]) -> np.ndarray:
        return draw_detections(frame, detections)
