"""Mint tests/golden/sinks.json + sinks_frame.npz from the REFERENCE's own result sinks.
TEST INFRASTRUCTURE ONLY; runs where /root/reference is mounted:  python -m oracle.make_sink_goldens

Inputs are a seeded result dictionary in the pipeline's schema (pipeliine.py:78-83,127-139): every vocabulary
character, float32-derived confidences, frames without detections, entries with missing keys.  Outputs are what
processing_service.py:59-137,188-218 and video_processing.py:169-216 return for it.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import reference_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHARS = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ "


def sample_results(seed: int = 0, frames: int = 6):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(frames):
        dets = []
        for _ in range(int(rng.integers(0, 5)) if i != 2 else 0):
            x1, y1 = int(rng.integers(0, 200)), int(rng.integers(30, 120))
            n = int(rng.integers(0, 20))
            dets.append({"bbox": [x1, y1, x1 + int(rng.integers(11, 100)), y1 + int(rng.integers(11, 40))],
                         "text": "".join(CHARS[int(k)] for k in rng.integers(0, len(CHARS), n)),
                         "detection_confidence": float(np.float32(rng.random())),
                         "recognition_confidence": float(np.float32(rng.random())),
                         "polygon": rng.integers(0, 640, (4, 2)).tolist()})
        out.append({"frame_number": i, "timestamp": i / 10.0, "detections": dets})
    out[1]["detections"].append({"bbox": [5, 40, 60, 70], "text": CHARS})          # missing confidences
    out.append({"detections": [{"text": "no bbox"}]})                              # missing frame keys and bbox
    summary = {"total_frames": frames + 1, "frames_with_text": 4, "total_detections": 11, "unique_texts": 2,
               "detected_texts": ["a<b", "R&D"], "avg_detection_confidence": 0.5, "avg_recognition_confidence": 0.25,
               "processing_time_seconds": 1.5, "fps_processed": 4.0, "note": ""}
    return {"status": "success", "results": out, "summary": summary,
            "video_info": {"fps": 30.0, "frame_count": 70, "width": 320, "height": 240}}


def overlay_detections(data):
    """Every detection of the sample (a missing confidence defaults to 0.0, :192) plus one malformed box (skipped, :194)."""
    return [d for f in data["results"] for d in f["detections"] if "bbox" in d] + [{"bbox": [1, 2, 3]}]


def main():
    to_csv, to_xml, draw, save = R.reference_sinks()
    data = sample_results()
    db = {}
    complete = {**data, "results": data["results"][:-1]}                           # the DB sink needs every key
    complete["results"][1] = {**complete["results"][1], "detections": complete["results"][1]["detections"][:-1]}
    save(db, 7, complete)
    frame = np.full((240, 320, 3), 90, np.uint8)
    drawn = draw(frame.copy(), overlay_detections(data))
    json.dump({"input": data, "csv": to_csv(data), "xml": to_xml(data), "empty_csv": to_csv({}), "empty_xml": to_xml({}),
               "db_input_video_id": 7, "db_frames": [dict(r) for r in db["frames"]],
               "db_detections": [dict(r) for r in db["detections"]], "db_frame_id_base": 1000},
              open(os.path.join(ROOT, "tests/golden/sinks.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/sinks_frame.npz"), drawn=drawn)
    print("wrote tests/golden/sinks.json, sinks_frame.npz")


if __name__ == "__main__":
    main()
