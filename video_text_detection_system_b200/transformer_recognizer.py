"""TransformerRecognizer with the reference's call surface (app/ml/models/text_recognizer.py:39-69): the TrOCR branch.

The reference builds `TrOCRProcessor` + `VisionEncoderDecoderModel` with `from_pretrained("microsoft/trocr-base-printed")`
(a download, text_recognizer.py:41-42) and greedy-generates at most 50 tokens per crop (:58).  There is no network
here and no checkpoint ships with the repository, so -- exactly like the reference offline -- construction raises unless
a local TrOCR checkpoint directory is named (argument `model_name` or the environment variable VTD_TROCR_DIR).  It never
falls back to the CRNN: a drop-in that silently swaps the recogniser is worse than one that refuses.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, List, Optional

import numpy as np

logger = logging.getLogger(__name__)


class TransformerRecognizer:
    def __init__(self, model_name: str = "microsoft/trocr-base-printed", dtype: Optional[str] = None):
        path = model_name if os.path.isdir(model_name) else os.environ.get("VTD_TROCR_DIR", "")
        if not path or not os.path.isdir(path):
            # what from_pretrained does offline (text_recognizer.py:41): OSError
            raise OSError("TrOCR checkpoint %r is not available locally (no network); set VTD_TROCR_DIR to a directory "
                          "holding the HuggingFace checkpoint, or construct TextRecognizer(use_transformer=False) for "
                          "the CRNN/CTC branch" % (model_name,))
        raise NotImplementedError("the sm_100a TrOCR branch is not built yet (SURVEY.md section 8f, N1)")

    def recognize(self, image: np.ndarray) -> Dict[str, Any]:      # pragma: no cover - unreachable until N1 lands
        return {"text": "", "confidence": 0.0}

    def recognize_batch(self, images: List[np.ndarray]) -> List[Dict[str, Any]]:      # pragma: no cover
        return [self.recognize(im) for im in images]
