"""Load the reference's hot-path source files by path.  TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference is mounted (the build container); the GPU box
does not have it, so nothing that runs there may call this.  Used by
oracle/make_goldens.py (to mint tests/golden/*.npz from the reference's own
code) and by tests/test_oracle_vs_reference.py (skipped when the tree is absent).

The package cannot be imported normally (D1-D4), so each file is
exec'd into a fresh module with the two missing typing names pre-seeded and the
removed `np.int0` alias restored.  No reference source is copied.
"""
from __future__ import annotations

import os
import types
import typing

import numpy as np
import torch
import torch.nn as nn
import torchvision

REF_ROOT = os.environ.get("VTD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "app/ml/models/text_detector.py"))


def _load(name: str, rel: str, seed=None):
    path = os.path.join(REF_ROOT, rel)
    m = types.ModuleType(name)
    m.__file__ = path
    m.__dict__.update(seed or {})
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), m.__dict__)
    return m


_cache = {}


def modules():
    """(text_detector module, text_recognizer module) of the reference."""
    if "m" not in _cache:
        if not hasattr(np, "int0"):
            np.int0 = np.intp                                          # D8
        det = _load("ref_text_detector", "app/ml/models/text_detector.py")
        rec = _load("ref_text_recognizer", "app/ml/models/text_recognizer.py",
                    {"Tuple": typing.Tuple})                          # D1
        _cache["m"] = (det, rec)
    return _cache["m"]


def reference_dbnet(backbone: str, state_dict) -> nn.Module:
    """A DBNet assembled from the REFERENCE's FeaturePyramidNetwork and DBHead classes
    and torchvision's ResNet exactly as text_detector.py:17-23 does, bypassing the
    constructor's download (D7) and its resnet50-only branch (D6)."""
    det, _ = modules()
    net = det.DBNet.__new__(det.DBNet)
    nn.Module.__init__(net)
    if backbone == "resnet50":
        rn, cin = torchvision.models.resnet50(weights=None), 2048
    else:
        rn, cin = torchvision.models.resnet18(weights=None), 512
    net.backbone = nn.Sequential(*list(rn.children())[:-2])
    net.fpn = det.FeaturePyramidNetwork(cin)
    net.head = det.DBHead(256)
    net.load_state_dict(state_dict, strict=True)
    return net.eval()


def reference_detector(net: nn.Module):
    """A reference TextDetector instance without running its __init__ (D7)."""
    import torchvision.transforms as T
    det, _ = modules()
    d = object.__new__(det.TextDetector)
    d.device = "cpu"
    d.model = net
    d.transform = T.Compose([T.ToPILImage(), T.Resize((640, 640)), T.ToTensor(),
                             T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return d


def reference_recognizer(state_dict=None):
    """The reference TextRecognizer(use_transformer=False), unmodified."""
    _, rec = modules()
    r = rec.TextRecognizer(use_transformer=False)
    r.device = "cpu"
    r.model.to("cpu")
    if state_dict is not None:
        r.model.load_state_dict(state_dict, strict=True)
    r.model.eval()
    return r


def reference_sinks():
    """The reference's result sinks, unmodified, as plain callables:
    (export_results_csv, export_results_xml, _draw_detections, save_results_to_database).
    Their modules import celery / sqlalchemy / settings, which are absent here, so only the four function bodies
    are compiled (by AST, from the files where they lie) into a namespace holding what they use.  The database sink's
    Pydantic/CRUD collaborators are replaced by recorders that return what they were called with."""
    import ast
    import asyncio
    import csv
    import io
    import logging
    import xml.etree.ElementTree as ET
    from typing import Any, Dict, List, Optional

    import cv2

    def functions(rel, names):
        path = os.path.join(REF_ROOT, rel)
        with open(path) as f:
            tree = ast.parse(f.read(), path)
        found = [n for n in ast.walk(tree)
                 if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)) and n.name in names]
        mod = ast.Module(body=found, type_ignores=[])
        return compile(mod, path, "exec")

    class Row(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    class FrameCRUD:
        @staticmethod
        def create_bulk(db, frames):
            db["frames"] = frames
            return [Row(id=1000 + i, **f) for i, f in enumerate(frames)]

    class TextDetectionCRUD:
        @staticmethod
        def create_bulk(db, detections):
            db["detections"] = detections
            return detections

    ns = {"io": io, "csv": csv, "ET": ET, "cv2": cv2, "np": np, "Dict": Dict, "Any": Any, "List": List,
          "Optional": Optional, "logger": logging.getLogger("reference_sinks"), "Session": dict,
          "FrameCreate": Row, "TextDetectionCreate": Row, "FrameCRUD": FrameCRUD,
          "TextDetectionCRUD": TextDetectionCRUD}
    exec(functions("app/services/processing_service.py",
                   {"export_results_csv", "export_results_xml", "_draw_detections"}), ns)
    exec(functions("app/tasks/video_processing.py", {"save_results_to_database"}), ns)

    def run(coro_fn):
        return lambda data: asyncio.run(coro_fn(None, data))

    return (run(ns["export_results_csv"]), run(ns["export_results_xml"]),
            lambda frame, dets: ns["_draw_detections"](None, frame, dets), ns["save_results_to_database"])
