"""The reference's own GPU path - stock PyTorch eager - as a timed BASELINE.  TEST / BENCH INFRASTRUCTURE.

SURVEY.md 2.1: on a CUDA machine the reference runs `transformers` DINOv3ViTModel under fp16 autocast with SDPA
attention (cbas.py:431-436, :672-677) and `classifier_head.ClassifierLSTMDeltas` with cuDNN's LSTM inside the
per-frame window loop of `infer_file` (cbas.py:497-551).  bench.py's `gpu_eager_baseline` leg times exactly that on the
same B200, next to the CPU baseline; nothing here is imported by the product (cbas_b200/).

  * encoder: the installed transformers implementation (the reference's third-party dependency, requirements.txt:26),
    random-init from a seed; the reference's preprocessing restated from cbas.py:431,672-675 (REFERENCE mode) or the
    HF processor arithmetic run with torch ops on the GPU (PROCESSOR mode, the 224-px configs);
  * head: EagerHead restates classifier_head.py:57-172 module for module (same parameter names, nn.LSTM -> cuDNN),
    because /root/reference does not exist on the GPU box; tests/test_oracle.py pins it to oracle.head.head_forward
    (itself pinned to the reference's outputs) and, when the checkout is present, to the reference module live.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoder as oenc


# ------------------------------------------------------------------------------------------------ encoder
@torch.no_grad()
def eager_encode_chunk_reference_loop(model, frames_u8: np.ndarray, device) -> np.ndarray:
    """One iteration of encode_file's chunk loop exactly as the reference runs it on CUDA (cbas.py:431-438 with
    DinoEncoder.forward, :672-677): float64 divide on the host, pageable H2D of the fp32 green plane, x3 channel
    repeat on the device, fp16 autocast forward, CLS row, synchronous .cpu()."""
    frames_tensor = torch.from_numpy(frames_u8[:, :, :, 1] / 255.0).float()
    with torch.autocast(device_type="cuda", enabled=True):
        x = frames_tensor.unsqueeze(1).to(device)                      # [B,1,H,W]
        B, S, H, W = x.shape
        x = x.unsqueeze(2).repeat(1, 1, 3, 1, 1).reshape(B * S, 3, H, W)
        out = model(x).last_hidden_state[:, 0, :]
        return out.float().cpu().numpy()


@torch.no_grad()
def eager_encode_processor_device(model, frames_dev_u8: torch.Tensor, size: int) -> torch.Tensor:
    """PROCESSOR-mode chunk with everything on the device (frames already resident, like bench.py's `value` leg):
    HF DINOv3ViTImageProcessor arithmetic (image_processing_dinov3_vit.py:45-86) with torch ops, then the fp16
    autocast forward and the CLS row."""
    x = frames_dev_u8.permute(0, 3, 1, 2).float() * (1.0 / 255.0)
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=True)
    mean = torch.tensor(oenc.IMAGENET_MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(oenc.IMAGENET_STD, device=x.device).view(1, 3, 1, 1)
    x = (x - mean) / std
    with torch.autocast(device_type="cuda", enabled=True):
        return model(x).last_hidden_state[:, 0, :].float()


def time_eager_encoder(arch: str, device, chunk: int, src_hw, size: int, steps: int, warmup: int = 2) -> Dict:
    """frames/s of the stock eager path on `device` for `chunk`-frame chunks: device-resident PROCESSOR mode (CUDA
    events, comparable with bench.py `value`) and the reference's literal host loop in REFERENCE mode at `size` px
    (wall clock, comparable with `e2e`)."""
    model = oenc.build_hf_model(arch, seed=0).to(device)
    g = torch.Generator(device=device).manual_seed(7)
    frames = [torch.randint(0, 256, (chunk, *src_hw, 3), dtype=torch.uint8, device=device, generator=g) for _ in range(2)]
    for i in range(warmup):
        eager_encode_processor_device(model, frames[i & 1], size)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = eager_encode_processor_device(model, frames[i & 1], size)
    e1.record()
    torch.cuda.synchronize(device)
    dev_ms = e0.elapsed_time(e1) / steps
    assert bool(torch.isfinite(out).all())
    del frames
    host = [np.random.default_rng(i).integers(0, 256, (chunk, size, size, 3), dtype=np.uint8) for i in range(2)]
    eager_encode_chunk_reference_loop(model, host[0], device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    n_loop = max(2, steps // 4)
    for i in range(n_loop):
        eager_encode_chunk_reference_loop(model, host[i & 1], device)
    torch.cuda.synchronize(device)
    loop_s = (time.perf_counter() - t0) / n_loop
    attn = getattr(model.config, "_attn_implementation", "?")
    return {"value": chunk / (dev_ms / 1000.0), "unit": "frames/s", "ms_per_step": dev_ms,
            "reference_loop": {"value": chunk / loop_s, "unit": "frames/s", "ms_per_step": loop_s * 1000.0,
                               "what": f"cbas.py:431-438 + DinoEncoder.forward literally: {size}x{size} frames, float64 /255 on "
                                       "the host, pageable H2D of fp32 planes, fp16 autocast, .cpu() per chunk"},
            "what": f"transformers DINOv3ViTModel ({arch}, random init) under torch.autocast('cuda') [fp16], attention "
                    f"'{attn}', {chunk}-frame chunks, frames resident on the device, HF-processor arithmetic as torch ops; "
                    "CUDA events", "steps": steps}


# ------------------------------------------------------------------------------------------------ head
class EagerHead(nn.Module):
    """classifier_head.py:57-172 restated module for module (eval-mode forward)."""

    def __init__(self, in_features, out_features, seq_len=31, bottleneck_dim=128, dropout_p=0.15, use_acceleration=True,
                 ema_alpha=0.3, center_window_size=5, lstm_hidden_size=64, lstm_layers=1):
        super().__init__()
        self.seq_len, self.sw, self.hsl = seq_len, center_window_size, seq_len // 2
        self.use_acceleration, self.ema_alpha = use_acceleration, ema_alpha
        mk = lambda: nn.Sequential(nn.Linear(in_features, bottleneck_dim), nn.GELU(), nn.Dropout(0.1))
        self.cls_bottleneck, self.delta_bottleneck = mk(), mk()
        self.cls_ln, self.delta_ln = nn.LayerNorm(bottleneck_dim), nn.LayerNorm(bottleneck_dim)
        if use_acceleration:
            self.acc_bottleneck, self.acc_ln = mk(), nn.LayerNorm(bottleneck_dim)
        self.lin0 = nn.Sequential(nn.Linear(bottleneck_dim * (3 if use_acceleration else 2), 256), nn.GELU(),
                                  nn.Dropout(dropout_p))
        self.gate = nn.Parameter(torch.tensor(0.2))
        self.attention_head = nn.Linear(lstm_hidden_size * 2, 1)
        self.attention_temp = nn.Parameter(torch.tensor(1.0))
        self.lin1 = nn.Linear(in_features, out_features)
        self.lin2 = nn.Linear(lstm_hidden_size * 2, out_features)
        self.lstm = nn.LSTM(256, lstm_hidden_size, num_layers=lstm_layers, batch_first=True, bidirectional=True)
        self.eval()

    def _deltas(self, x):  # classifier_head.py:102-117
        s = torch.zeros_like(x)
        s[:, 0] = x[:, 0]
        for t in range(1, x.shape[1]):
            s[:, t] = torch.lerp(s[:, t - 1], x[:, t], self.ema_alpha)
        mode = "reflect" if x.shape[1] >= 3 else "replicate"
        padded = F.pad(s.permute(0, 2, 1), (2, 0), mode).permute(0, 2, 1)
        dx = padded[:, 1:] - padded[:, :-1]
        return s, dx[:, 1:], dx[:, 1:] - dx[:, :-1]

    def forward(self, x):  # classifier_head.py:119-172
        cls_s, d_s, a_s = self._deltas(x.float())
        L = x.size(1)
        l, r = max(0, self.hsl - self.sw), min(L, self.hsl + self.sw + 1)
        idx = min(max(0, L // 2), L - 1)
        linear_logits = self.lin1(cls_s[:, idx]) if l >= r else self.lin1(cls_s[:, l:r]).mean(dim=1)
        parts = [self.cls_ln(self.cls_bottleneck(cls_s)), self.delta_ln(self.delta_bottleneck(d_s))]
        if self.use_acceleration:
            parts.append(self.acc_ln(self.acc_bottleneck(a_s)))
        z = self.lin0(torch.cat(parts, dim=-1))
        z = z - z.mean(dim=1, keepdim=True, dtype=torch.float32)
        out, _ = self.lstm(z)
        if l >= r:
            rawm = out[:, idx]
        else:
            win = out[:, l:r]
            scores = self.attention_head(win).squeeze(-1) / (F.softplus(self.attention_temp) + 1e-3)
            rawm = (torch.softmax(scores, dim=1).unsqueeze(-1) * win).sum(dim=1)
        return torch.lerp(linear_logits, self.lin2(rawm), torch.sigmoid(self.gate)), rawm


def eager_head_from_state(sd: Dict[str, torch.Tensor], in_features: int, out_features: int, seq_len: int = 31) -> EagerHead:
    hs = sd["lstm.weight_hh_l0"].shape[1]
    layers = sum(1 for k in sd if k.startswith("lstm.weight_ih_l") and not k.endswith("_reverse"))
    head = EagerHead(in_features, out_features, seq_len=seq_len, bottleneck_dim=sd["cls_ln.weight"].shape[0],
                     use_acceleration="acc_ln.weight" in sd, lstm_hidden_size=hs, lstm_layers=layers)
    head.load_state_dict(sd, strict=True)
    return head.eval()


@torch.no_grad()
def eager_infer_loop(model: nn.Module, emb_f16: np.ndarray, seq_len: int, device, temperature: float = 1.0) -> np.ndarray:
    """infer_file's numeric loop exactly as the reference runs it (cbas.py:497-551): 20 000-frame chunks with context,
    replicate padding, one Python slice per frame, batches of 512 stacked on the host, H2D, forward, softmax, .cpu()."""
    total, half = emb_f16.shape[0], seq_len // 2
    all_probs = []
    for start in range(0, total, 20000):
        end = min(start + 20000, total)
        rs, re_ = max(0, start - half), min(total, end + half)
        ct = torch.from_numpy(emb_f16[rs:re_]).float()
        if start < half and half - start > 0:
            ct = torch.cat([ct[0:1].repeat(half - start, 1), ct], dim=0)
        if end > total - half and half - (total - end) > 0:
            ct = torch.cat([ct, ct[-1:].repeat(half - (total - end), 1)], dim=0)
        buf, n_t = [], end - start
        for i in range(n_t):
            buf.append(ct[i:i + seq_len])
            if len(buf) >= 512 or i == n_t - 1:
                logits, _ = model(torch.stack(buf).to(device))
                all_probs.extend(torch.softmax(logits / max(1e-3, temperature), dim=1).cpu().numpy())
                buf = []
    return np.array(all_probs)


def time_eager_head(sd: Dict[str, torch.Tensor], in_features: int, out_features: int, device, frames: int,
                    seq_len: int = 31) -> Dict:
    head = eager_head_from_state(sd, in_features, out_features, seq_len).to(device)
    emb = np.random.default_rng(0).standard_normal((frames, in_features)).astype(np.float16)
    eager_infer_loop(head, emb[:2048], seq_len, device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    probs = eager_infer_loop(head, emb, seq_len, device)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    assert probs.shape == (frames, out_features) and np.isfinite(probs).all()
    return {"value": frames / dt, "unit": "frames/s", "seconds": dt, "frames": frames,
            "what": "classifier_head.ClassifierLSTMDeltas restated module for module (nn.LSTM -> cuDNN) driven by "
                    "infer_file's window loop (cbas.py:497-551): per-frame Python slicing, 512-window batches stacked on "
                    "the host, H2D, fp32 forward, .cpu() per batch; wall clock"}
