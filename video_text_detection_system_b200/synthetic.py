"""Synthetic workloads of the benchmark: seeded frames, random-init weights of the reference
architecture with non-trivial BatchNorm statistics, and config 3's planted logit plane.  Input generators only -- no
reference arithmetic lives here; the CPU oracle (oracle/port.py) and bench.py both draw their inputs from this module so
that the B200 arm never imports the oracle."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn


def synthetic_frames(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """n uniform-random BGR frames [n,h,w,3] u8."""
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


def randomize_bn(module: nn.Module, seed: int) -> None:
    """Parity value distribution of SURVEY.md 8d: default-init BN is an identity and would hide folding bugs, so give
    every BN non-trivial statistics (running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5), weight ~ U(0.5,1.5),
    bias ~ N(0,0.1))."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)


def random_state_dicts(seed: int = 0, backbone: str = "resnet18") -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(detector, recogniser) state dicts in the reference's key layout: PyTorch default / torchvision-ResNet
    initialisation under `seed`, BatchNorm statistics randomised.  Built from the package's own parameter containers
    (models.DBNet / models.CRNN), no checkpoint and no download."""
    from .models import CRNN, DBNet
    torch.manual_seed(seed)
    det = DBNet(backbone, pretrained=False)
    randomize_bn(det, seed + 1000)
    torch.manual_seed(seed)
    rec = CRNN(97)
    randomize_bn(rec, seed + 1000)
    clone = lambda sd: {k: v.detach().clone() for k, v in sd.items()}
    return clone(det.state_dict()), clone(rec.state_dict())


def planted_logit_bias(n: int, det_h: int, det_w: int, seed: int = 0, boxes: int = 50,
                       inside: float = 8.0, outside: float = -8.0) -> np.ndarray:
    """Config 3's planted logit plane: `boxes` rotated rectangles per frame on a jittered 10x5 grid, w~U[60,100],
    h~U[20,36] detector pixels, angle~U[-15,15] degrees.  Returns [n, det_h, det_w] fp32 holding `inside` within the
    rectangles and `outside` elsewhere; added to the probability head's pre-sigmoid logit."""
    import cv2
    rng = np.random.default_rng(seed)
    out = np.full((n, det_h, det_w), outside, np.float32)
    gx, gy = 10, 5
    cw, ch = det_w / gx, det_h / gy
    for f in range(n):
        m = np.zeros((det_h, det_w), np.uint8)
        k = 0
        for j in range(gy):
            for i in range(gx):
                if k >= boxes:
                    break
                w = rng.uniform(60, 100)
                h = rng.uniform(20, 36)
                a = rng.uniform(-15, 15)
                cx = (i + 0.5) * cw + rng.uniform(-8, 8)
                cy = (j + 0.5) * ch + rng.uniform(-8, 8)
                pts = cv2.boxPoints(((float(cx), float(cy)), (float(w), float(h)), float(a)))
                cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
                k += 1
        out[f][m > 0] = inside
    return out


def random_trocr_model(kind: str = "base", seed: int = 0):
    """A randomly initialised HuggingFace VisionEncoderDecoderModel in the architecture of the reference's TrOCR branch
    (text_recognizer.py:41-42; the checkpoint itself is a download and unavailable offline).  kind "base" = the published
    configuration of microsoft/trocr-base-printed (ViT-B/16 @384, 12 + 12 layers, 341 M parameters); "tiny" = the same
    architecture shrunk (64x64 images, width 128, 2 + 2 layers) for quick checks.  Weight generator only: its state dict
    feeds Engine.load_trocr, and the CPU oracle (oracle/trocr_port.py) runs the very same module."""
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
