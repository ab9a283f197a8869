"""TransformerRecognizer with the reference's call surface (app/ml/models/text_recognizer.py:39-69): the TrOCR branch.

The reference builds `TrOCRProcessor` + `VisionEncoderDecoderModel` with `from_pretrained("microsoft/trocr-base-printed")`
(a download, text_recognizer.py:41-42) and greedy-generates at most 50 tokens per crop (:58); its "confidence" is the
constant 0.95 (:64).  Here the model runs inside libvtd_b200.so (csrc/trocr.cu, csrc/trocr_host.inc: ViT encoder and TrOCR
decoder on the tcgen05 GEMM kernels, flash attention, KV-cache decode, greedy argmax; the processor's 384x384 bilinear
resize + normalisation on the device); `transformers` is used only to READ a checkpoint directory and for the tokenizer
that turns token ids into text.

There is no network here and no checkpoint ships with the repository, so -- exactly like the reference offline --
construction raises unless a local TrOCR checkpoint directory is named (argument `model_name`, or the environment variable
VTD_TROCR_DIR) or a state dict is handed over.  It never falls back to the CRNN: a drop-in that silently swaps the
recogniser is worse than one that refuses.
"""
from __future__ import annotations

import logging
import os
import threading
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from ._lib import Engine

logger = logging.getLogger(__name__)


class TransformerRecognizer:
    def __init__(self, model_name: str = "microsoft/trocr-base-printed", dtype: Optional[str] = None, *, state_dict=None,
                 decode: Optional[Callable[[List[int]], str]] = None, crops_per_chunk: int = 64, max_length: int = 50,
                 device: int = 0):
        self.device = "cuda"
        self.max_length = int(max_length)
        self.crops_per_chunk = int(crops_per_chunk)
        self._dtype = (dtype or "fp16").lower()
        if self._dtype == "fp32":
            self._dtype = "fp16"                      # the transformer recogniser exists in the 16-bit speed tier only
        self._device_index = int(device)
        self._decode = decode
        self._lock = threading.Lock()
        self._engine: Optional[Engine] = None
        self.processor = None
        if state_dict is None:
            path = model_name if os.path.isdir(model_name) else os.environ.get("VTD_TROCR_DIR", "")
            if not path or not os.path.isdir(path):
                # what from_pretrained does offline (text_recognizer.py:41): OSError
                raise OSError("TrOCR checkpoint %r is not available locally (no network); set VTD_TROCR_DIR to a directory "
                              "holding the HuggingFace checkpoint, or construct TextRecognizer(use_transformer=False) for "
                              "the CRNN/CTC branch" % (model_name,))
            from transformers import TrOCRProcessor, VisionEncoderDecoderModel
            self.processor = TrOCRProcessor.from_pretrained(path, local_files_only=True)
            hf = VisionEncoderDecoderModel.from_pretrained(path, local_files_only=True)
            state_dict = hf.state_dict()
            cfg = hf.config
            self.special_ids = {cfg.decoder_start_token_id, cfg.pad_token_id, cfg.eos_token_id, getattr(cfg.decoder, "bos_token_id", 0)}
        else:
            self.special_ids = {0, 1, 2}
        self._state_dict = {k: v for k, v in state_dict.items()}
        self.model = self                              # the reference's attribute (recognizer.model.recognize)

    # ---- engine (lazy: CUDA is not touched before the first call, prefork-safe)
    def _eng(self) -> Engine:
        if self._engine is None:
            eng = Engine(device=self._device_index, dtype=self._dtype, det_h=32, det_w=32, max_batch=1, max_boxes=64, max_src_h=32,
                         max_src_w=32)
            eng.load_trocr(self._state_dict, crops_per_chunk=self.crops_per_chunk)
            self._engine = eng
        return self._engine

    def _text(self, ids: List[int]) -> str:
        if self.processor is not None:                 # processor.batch_decode(generated_ids, skip_special_tokens=True)[0] (:60)
            return self.processor.batch_decode([ids], skip_special_tokens=True)[0]
        if self._decode is not None:
            return self._decode(ids)
        return " ".join(str(i) for i in ids if i not in self.special_ids)

    def generate_ids(self, images: List[np.ndarray]) -> List[List[int]]:
        for im in images:
            if not isinstance(im, np.ndarray) or im.ndim != 3 or im.shape[2] != 3 or im.dtype != np.uint8 or im.size == 0:
                raise ValueError("expected non-empty HxWx3 uint8 BGR crops")
        with self._lock:
            ids, lens = self._eng().trocr_generate_crops(list(images), self.max_length)
        return [ids[i, :lens[i]].tolist() for i in range(len(images))]

    # ---- reference surface
    def recognize(self, image: np.ndarray) -> Dict[str, Any]:
        try:
            ids = self.generate_ids([image])[0]
            return {"text": self._text(ids), "confidence": 0.95}                   # text_recognizer.py:62-65
        except Exception as e:
            logger.error(f"Text recognition failed: {e}")
            return {"text": "", "confidence": 0.0}

    def recognize_batch(self, images: List[np.ndarray]) -> List[Dict[str, Any]]:
        """One device pass over all crops (the reference loops recognize() per crop, :103-104); a crop the reference would
        fail on individually fails individually here as well."""
        try:
            all_ids = self.generate_ids(list(images))
            return [{"text": self._text(ids), "confidence": 0.95} for ids in all_ids]
        except Exception:
            return [self.recognize(im) for im in images]
