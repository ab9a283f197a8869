"""CPU oracle of the transformer recogniser (TrOCR branch).  TEST INFRASTRUCTURE ONLY (see oracle/port.py).

The reference's TransformerRecognizer (app/ml/models/text_recognizer.py:39-69) is TrOCRProcessor +
VisionEncoderDecoderModel from the third-party `transformers` package (pinned 4.36.0 in requirements.txt:13; 5.5.0 is
what this image holds) with the checkpoint "microsoft/trocr-base-printed", which cannot be downloaded here.  Parity is
therefore pinned the only way it can be offline ("random-init-from-config parity only"): the SAME
HuggingFace classes are instantiated from the checkpoint's published configuration with seeded random weights, and the
CUDA path must reproduce their encoder states, their teacher-forced logits and their greedy `generate(max_length=50)`
on the same inputs.  The image processor is the checkpoint's (resize 384x384 PIL bilinear, rescale 1/255, mean/std 0.5).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


def build(kind: str = "base", seed: int = 0):
    """The HuggingFace module itself (seeded random weights in the checkpoint's configuration): the weights come from the
    package's input generator (synthetic.random_trocr_model), as for the other models; the arithmetic checked against is
    transformers' own forward / generate."""
    from video_text_detection_system_b200.synthetic import random_trocr_model
    return random_trocr_model(kind, seed)


def image_size(model) -> int:
    return int(model.config.encoder.image_size)


def processor_pixel_values(crops_bgr: Sequence[np.ndarray], size: int) -> torch.Tensor:
    """TransformerRecognizer.recognize's preprocessing (text_recognizer.py:48-56): BGR->RGB, PIL image, the checkpoint's
    image processor: resize (size, size) PIL bilinear, rescale 1/255, normalise mean 0.5 / std 0.5."""
    import cv2
    from PIL import Image
    out = []
    for im in crops_bgr:
        rgb = cv2.cvtColor(im, cv2.COLOR_BGR2RGB)
        pil = Image.fromarray(rgb).resize((size, size), resample=Image.BILINEAR)
        x = np.asarray(pil).astype(np.float32) * np.float32(1.0 / 255.0)
        x = (x - np.float32(0.5)) / np.float32(0.5)
        out.append(torch.from_numpy(x.transpose(2, 0, 1).copy()))
    return torch.stack(out)


def generate(model, pixel_values: torch.Tensor, max_length: int = 50) -> np.ndarray:
    """generate(pixel_values, max_length=50) as the reference calls it (:58): greedy; padded to max_length with pad."""
    with torch.no_grad():
        ids = model.generate(pixel_values, max_length=max_length, do_sample=False, num_beams=1)
    out = np.full((ids.shape[0], max_length), model.config.pad_token_id, np.int32)
    out[:, :ids.shape[1]] = ids.numpy()
    return out


def forward_logits(model, pixel_values: torch.Tensor, decoder_ids: np.ndarray):
    with torch.no_grad():
        o = model(pixel_values=pixel_values, decoder_input_ids=torch.from_numpy(decoder_ids).long(), output_hidden_states=False)
    return o.encoder_last_hidden_state.numpy(), o.logits.numpy()
