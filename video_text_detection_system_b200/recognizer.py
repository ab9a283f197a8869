"""TextRecognizer with the reference's call surface (app/ml/models/text_recognizer.py:71-167).

CRNN branch: crop resize -> conv stack -> 2-layer BiLSTM -> Linear -> softmax -> greedy decode, all in
libvtd_b200.so.  The TrOCR branch (text_recognizer.py:39-69) lives in
transformer_recognizer.py; like the reference's, its constructor raises when no TrOCR weights can be found -- it never
substitutes the CRNN.
"""
from __future__ import annotations

import logging
import os
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from ._lib import CHARS, ids_to_text
from .models import CRNN

logger = logging.getLogger(__name__)


class TextRecognizer:
    def __init__(self, model_path: str = None, use_transformer: bool = True, *, crop_w: int = 128,
                 dtype: Optional[str] = None, trocr_state_dict=None, trocr_decode=None, trocr_dir: Optional[str] = None):
        self.use_transformer = use_transformer
        self.device = "cuda"
        self.crop_w = int(crop_w)
        self._lock = threading.Lock()
        if use_transformer:
            # text_recognizer.py:76-77: the TrOCR branch.  As in the reference, construction fails when the weights are
            # not to be had (from_pretrained raises offline); no silent substitution of the CRNN.
            from .transformer_recognizer import TransformerRecognizer
            self.model = TransformerRecognizer(trocr_dir or "microsoft/trocr-base-printed", dtype=dtype,
                                               state_dict=trocr_state_dict, decode=trocr_decode)
            return
        self.vocab = self._build_vocab()
        self.model = CRNN(len(self.vocab))
        self.model.dtype_tier = (dtype or os.environ.get("VTD_DTYPE", "fp16")).lower()
        if model_path:
            self.load_model(model_path)
        self.model.eval()

    def _build_vocab(self) -> Dict[str, int]:
        vocab = {char: i + 1 for i, char in enumerate(CHARS)}          # text_recognizer.py:86-91
        vocab["<blank>"] = 0
        vocab["<unk>"] = len(vocab)
        return vocab

    def load_model(self, model_path: str):
        try:
            checkpoint = torch.load(model_path, map_location="cpu")
            self.model.load_state_dict(checkpoint["model_state_dict"])
            logger.info(f"CRNN model loaded from {model_path}")
        except Exception as e:
            logger.error(f"Failed to load CRNN model: {e}")
            raise

    def _forward_is_patched(self) -> bool:
        return (not self.use_transformer) and "forward" in vars(self.model)

    def recognize_batch(self, images: List[np.ndarray]) -> List[Dict[str, Any]]:
        if self.use_transformer:                                          # text_recognizer.py:103-104
            return self.model.recognize_batch(images)
        return self._recognize_crnn_batch(images)

    def recognize(self, image: np.ndarray) -> Dict[str, Any]:
        if self.use_transformer:                                          # text_recognizer.py:109-110
            return self.model.recognize(image)
        return self._recognize_crnn_batch([image])[0]

    def _recognize_crnn_batch(self, images: List[np.ndarray]) -> List[Dict[str, Any]]:
        try:
            if len(images) == 0:
                raise ValueError("empty batch")                           # torch.stack([]) raises in the reference
            eng = self.model.get_engine(self.crop_w)
            if self._forward_is_patched():
                # test-compat path (tests/test_models.py:73): model.forward was replaced by the caller
                with self._lock:
                    eng.recognize_crops(list(images))                     # fills the resized inputs
                    x = torch.from_numpy(eng.debug_tensor("crops", min(len(images), 1024)))
                with torch.no_grad():
                    outputs = self.model(x)
                    predictions = torch.softmax(outputs, dim=2)
                return [dict(zip(("text", "confidence"), self._decode_prediction(p))) for p in predictions]
            with self._lock:
                ids, lens, conf, _ = eng.recognize_crops(list(images))
            return [{"text": ids_to_text(ids[i, :lens[i]].tolist()), "confidence": float(conf[i])}
                    for i in range(len(images))]
        except Exception as e:
            logger.error(f"CRNN batch recognition failed: {e}")
            return [{"text": "", "confidence": 0.0}] * len(images)

    def _decode_prediction(self, prediction: torch.Tensor) -> Tuple[str, float]:
        """text_recognizer.py:142-167 on one [T,V] matrix of probabilities (reference semantics, GPU decode)."""
        p = prediction.detach().float().cpu().numpy()
        eng = self.model.get_engine(self.crop_w)
        with self._lock:
            ids, lens, conf = eng.ctc_decode(p, is_prob=True)
        return ids_to_text(ids[0, :lens[0]].tolist()), float(conf[0])
