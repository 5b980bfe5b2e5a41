"""CPU oracle for the streamed encoder.  TEST INFRASTRUCTURE.

The reference's ViT arithmetic is third-party: cbas.py:657 loads `transformers.AutoModel`, cbas.py:676 calls it.
The same installed implementation (transformers.models.dinov3_vit, v5.5.0 here) is therefore the oracle for the
forward pass; this module restates only what the reference wraps around it:
  * cbas.py:431        frames_np[:, :, :, 1] / 255.0  (float64 divide, then .float())
  * cbas.py:672-675    unsqueeze / repeat x3 / reshape to (B*S, 3, H, W)
  * cbas.py:677        last_hidden_state[:, 0, :]  (CLS after the final norm), with 768 -> config.hidden_size
  * PROCESSOR mode     transformers DINOv3ViTImageProcessor._preprocess (image_processing_dinov3_vit.py:45-86):
                       rescale 1/255 -> tvF.resize(bilinear, antialias=True) -> normalize(ImageNet mean/std)
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

ARCH = {
    "vits16": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536),
    "vitb16": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072),
    "vitl16": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
}


def build_hf_model(arch: str = "vitb16", seed: int = 0, init_scale: float = 1.0, **overrides):
    """Random-init DINOv3ViTModel of the named architecture (4 register tokens, defaults of
    configuration_dinov3_vit.py:74-101).  init_scale > 1 multiplies every Linear/Conv weight after init so the
    embeddings depend visibly on the input (default init makes all frames' CLS nearly parallel)."""
    from transformers import DINOv3ViTConfig, DINOv3ViTModel
    kw = dict(ARCH[arch], num_register_tokens=4)
    kw.update(overrides)
    torch.manual_seed(seed)
    model = DINOv3ViTModel(DINOv3ViTConfig(**kw)).eval()
    if init_scale != 1.0:
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("weight") and p.dim() >= 2:
                    p.mul_(init_scale)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


ARCH_DINOV2 = {
    "dinov2reg-s14": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6),
    "dinov2reg-b14": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12),
}


def build_hf_dinov2_model(arch: str = "dinov2reg-b14", seed: int = 0, init_scale: float = 1.0, **overrides):
    """Random-init transformers Dinov2WithRegistersModel with the geometry of facebook/dinov2-with-registers-*
    (patch 14, image_size 518 => 37x37 learned position grid, 4 register tokens) - CBAS's default encoder
    (cbas.py:1030-1033).  The reference calls it exactly like the DINOv3 model (cbas.py:672-677)."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    kw = dict(ARCH_DINOV2[arch], patch_size=14, image_size=518, num_register_tokens=4)
    kw.update(overrides)
    torch.manual_seed(seed)
    model = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(**kw)).eval()
    with torch.no_grad():
        # the default init leaves register tokens at zero and position embeddings at N(0,1): keep the latter small
        # enough that the patch content still matters, and give the registers something to carry
        model.embeddings.register_tokens.normal_(0.0, 0.02)
        model.embeddings.position_embeddings.mul_(0.02)
        model.embeddings.cls_token.mul_(0.02)
        if init_scale != 1.0:
            for name, p in model.named_parameters():
                if name.endswith("weight") and p.dim() >= 2:
                    p.mul_(init_scale)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def preprocess_reference(frames_u8: np.ndarray) -> torch.Tensor:
    """(n,H,W,3) uint8 RGB -> (n,3,H,W) float32: green/255 replicated (cbas.py:431,672-675)."""
    x = torch.from_numpy(frames_u8[:, :, :, 1] / 255.0).float()
    return x.unsqueeze(1).repeat(1, 3, 1, 1)


def preprocess_processor(frames_u8: np.ndarray, size: int = 224) -> torch.Tensor:
    """(n,H,W,3) uint8 RGB -> (n,3,size,size) float32 the way DINOv3ViTImageProcessor does it."""
    x = torch.from_numpy(frames_u8).permute(0, 3, 1, 2).float() * (1.0 / 255.0)
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=True)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


@torch.no_grad()
def encode(model, frames_u8: np.ndarray, mode: str = "reference", size: int = 224, batch: int = 32) -> np.ndarray:
    """CLS embeddings [n, D] float32 for uint8 frames, fp32 on CPU (autocast disabled, cbas.py:434)."""
    outs = []
    for i in range(0, len(frames_u8), batch):
        chunk = frames_u8[i:i + batch]
        x = preprocess_reference(chunk) if mode == "reference" else preprocess_processor(chunk, size)
        outs.append(model(x).last_hidden_state[:, 0, :].float().numpy())
    return np.concatenate(outs) if outs else np.zeros((0, model.config.hidden_size), np.float32)


@torch.no_grad()
def hidden_states(model, pixel_values: torch.Tensor) -> List[torch.Tensor]:
    """Residual stream after the embeddings and after each block (before the final norm):
    the per-layer taps the GPU parity test compares against."""
    emb = model.embeddings(pixel_values)
    if not hasattr(model, "rope_embeddings"):  # Dinov2WithRegistersModel: absolute positions, plain layer stack
        hs, h = [emb], emb
        for layer in model.encoder.layer:
            h = layer(h)
            h = h[0] if isinstance(h, tuple) else h
            hs.append(h)
        return hs
    pos = model.rope_embeddings(pixel_values)
    hs = [emb]
    h = emb
    for layer in model.model.layer:
        h = layer(h, position_embeddings=pos)
        hs.append(h)
    return hs


def synthetic_frames(n: int, h: int, w: int, seed: int = 0, structured: bool = True) -> np.ndarray:
    """Seeded uint8 RGB frames.  structured: gradient background + moving Gaussian blob + per-frame noise
    (input-dependent embeddings); otherwise uniform noise."""
    rng = np.random.default_rng(seed)
    if not structured:
        return rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        cx, cy = (0.2 + 0.6 * ((i * 0.37) % 1.0)) * w, (0.3 + 0.4 * ((i * 0.61) % 1.0)) * h
        blob = 160.0 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2.0 * (0.08 * w + 2 * (i % 5)) ** 2))
        for c in range(3):
            bg = 40.0 + 60.0 * (xx / w if c != 1 else yy / h) + 10.0 * c
            noise = rng.normal(0.0, 12.0, (h, w))
            out[i, :, :, c] = np.clip(bg + blob * (0.6 + 0.2 * c) + noise, 0, 255).astype(np.uint8)
    return out
