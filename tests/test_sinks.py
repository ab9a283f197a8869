"""CPU: the result sinks (host half) against goldens minted from the reference's own functions
(oracle/make_sink_goldens.py), against the reference itself where it is mounted, and against csv / xml.etree on
hostile strings."""
import asyncio
import csv
import io
import json
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from oracle import reference_loader
from video_text_detection_system_b200 import sinks

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g():
    return json.load(open(os.path.join(GOLDEN, "sinks.json")))


def test_csv_and_xml_match_reference_goldens(g):
    assert sinks.export_results_csv(g["input"]) == g["csv"]
    assert sinks.export_results_xml(g["input"]) == g["xml"]
    assert sinks.export_results_csv({}) == g["empty_csv"]
    assert sinks.export_results_xml({}) == g["empty_xml"]
    rows = list(csv.reader(io.StringIO(g["csv"])))
    assert tuple(rows[0]) == sinks.CSV_HEADER
    assert len(rows) - 1 == sum(len(f["detections"]) for f in g["input"]["results"])
    root = ET.fromstring(g["xml"])
    assert len(root.find("frames")) == len(g["input"]["results"])


def test_database_rows_match_reference_goldens(g):
    data = dict(g["input"])
    data["results"] = [dict(f) for f in data["results"][:-1]]
    data["results"][1]["detections"] = data["results"][1]["detections"][:-1]
    frames, dets = sinks.database_rows(g["db_input_video_id"], data)
    assert frames == g["db_frames"]
    ids = {f["frame_number"]: g["db_frame_id_base"] + i for i, f in enumerate(frames)}   # frame_mapping, :187-188
    mapped = [{"frame_id": ids[d["frame_number"]], **{k: v for k, v in d.items() if k != "frame_number"}} for d in dets]
    assert mapped == g["db_detections"]
    with pytest.raises(KeyError):
        sinks.database_rows(1, g["input"])                 # the incomplete entries raise, as the reference (:208-210)


def test_overlay_matches_reference_golden(g):
    want = np.load(os.path.join(GOLDEN, "sinks_frame.npz"))["drawn"]
    frame = np.full((240, 320, 3), 90, np.uint8)
    dets = [d for f in g["input"]["results"] for d in f["detections"] if "bbox" in d] + [{"bbox": [1, 2, 3]}]
    out = sinks.draw_detections(frame, dets)
    assert out is frame and np.array_equal(out, want) and (want != 90).any()


def test_xml_serialisation_equals_elementtree_on_hostile_strings():
    texts = ['a&b<c>d"e\'f', "tab\there", "line\nbreak\r", "", " ", "]]>", "&amp;", "ünï©ode ✓", "<Point x=\"1\" />"]
    data = {"summary": {"k": "<&>", "empty": "", "n": 3, "lst": ["a<b", "c&d"]},
            "results": [{"frame_number": i, "timestamp": 0.1 * i,
                         "detections": [{"bbox": [i, -i, 10 * i, 7], "text": t, "detection_confidence": 1e-7,
                                         "recognition_confidence": 1.0}]} for i, t in enumerate(texts)]}
    root = ET.Element("video_text_detection")
    s = ET.SubElement(root, "summary")
    for k, v in data["summary"].items():
        ET.SubElement(s, k).text = str(v)
    fs = ET.SubElement(root, "frames")
    for fr in data["results"]:
        f = ET.SubElement(fs, "frame", number=str(fr["frame_number"]), timestamp=str(fr["timestamp"]))
        for d in fr["detections"]:
            o = ET.SubElement(f, "object", transcription=d["text"], detection_confidence=str(d["detection_confidence"]),
                              recognition_confidence=str(d["recognition_confidence"]))
            b = d["bbox"]
            for x, y in ((b[0], b[1]), (b[2], b[1]), (b[2], b[3]), (b[0], b[3])):
                ET.SubElement(o, "Point", x=str(x), y=str(y))
    assert sinks.export_results_xml(data) == ET.tostring(root, encoding="unicode")


def test_error_convention_and_method_names():
    assert sinks.export_results_csv({"results": 5}) == ""          # swallowed + logged, '' (:86-88, :135-137)
    assert sinks.export_results_xml({"results": 5}) == ""
    rs = sinks.ResultSinks()
    data = {"results": [{"frame_number": 1, "timestamp": 0.5, "detections": [{"bbox": [1, 2, 3, 4], "text": "x"}]}]}
    assert asyncio.run(rs.export_results_csv(data)) == sinks.export_results_csv(data)
    assert asyncio.run(rs.export_results_xml(data)) == sinks.export_results_xml(data)
    assert rs._draw_detections(np.zeros((8, 8, 3), np.uint8), []).shape == (8, 8, 3)


@pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not mounted")
def test_sinks_equal_reference_live(g):
    to_csv, to_xml, draw, save = reference_loader.reference_sinks()
    from oracle.make_sink_goldens import sample_results
    for seed in (1, 2, 3):
        data = sample_results(seed, frames=9)
        assert sinks.export_results_csv(data) == to_csv(data)
        assert sinks.export_results_xml(data) == to_xml(data)
        dets = [d for f in data["results"] for d in f["detections"] if "bbox" in d]
        a = draw(np.full((240, 320, 3), 30, np.uint8), dets)
        b = sinks.draw_detections(np.full((240, 320, 3), 30, np.uint8), dets)
        assert np.array_equal(a, b)
