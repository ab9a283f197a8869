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
    """kind "base" = the configuration of microsoft/trocr-base-printed (ViT-B/16 @384, 12+12 layers, 341 M parameters);
    "tiny" = the same architecture shrunk (64x64 images, width 128, 2+2 layers) for quick checks."""
    from transformers import TrOCRConfig, ViTConfig, VisionEncoderDecoderConfig, VisionEncoderDecoderModel
    if kind == "base":
        enc = ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=384,
                        patch_size=16, qkv_bias=False, hidden_act="gelu", layer_norm_eps=1e-12)
        dec = TrOCRConfig(vocab_size=50265, d_model=1024, decoder_layers=12, decoder_attention_heads=16, decoder_ffn_dim=4096,
                          activation_function="gelu", max_position_embeddings=512, scale_embedding=False,
                          use_learned_position_embeddings=True, layernorm_embedding=True, cross_attention_hidden_size=768)
    else:
        enc = ViTConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=64,
                        patch_size=16, qkv_bias=False, hidden_act="gelu", layer_norm_eps=1e-12)
        dec = TrOCRConfig(vocab_size=300, d_model=128, decoder_layers=2, decoder_attention_heads=2, decoder_ffn_dim=256,
                          activation_function="gelu", max_position_embeddings=64, scale_embedding=False,
                          use_learned_position_embeddings=True, layernorm_embedding=True, cross_attention_hidden_size=128)
    cfg = VisionEncoderDecoderConfig.from_encoder_decoder_configs(enc, dec)
    cfg.decoder_start_token_id, cfg.pad_token_id, cfg.eos_token_id = 2, 1, 2
    torch.manual_seed(seed)
    model = VisionEncoderDecoderModel(cfg).eval()
    # random-init LayerNorms are identities and the default 0.02 init leaves every logit near zero: give the norms
    # non-trivial affine parameters and widen the output projection so that greedy decoding is decided by real margins
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layernorm" in n.lower() or "layer_norm" in n.lower():
                if n.endswith("weight"):
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        model.decoder.output_projection.weight.mul_(8.0)
        if kind != "base":
            model.decoder.output_projection.weight[cfg.eos_token_id].zero_()     # the tiny net would emit EOS at once
    return model


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
