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
