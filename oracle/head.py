"""CPU restatement of the CBAS temporal-delta BiLSTM head and of infer_file's window loop.  TEST INFRASTRUCTURE.

Follows /root/reference/backend/classifier_head.py:57-172 (ClassifierLSTMDeltas) and
/root/reference/backend/cbas.py:481-551 (infer_file: chunking, replicate padding, batches of 512, temperature
softmax).  Written against plain tensors (no nn.Module, no nn.LSTM) so every step is explicit; pinned against
the reference module's own outputs by tests/golden/head_*.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HEAD_KEYS_1LAYER = (
    "gate", "attention_temp",
    "cls_bottleneck.0.weight", "cls_bottleneck.0.bias", "delta_bottleneck.0.weight", "delta_bottleneck.0.bias",
    "acc_bottleneck.0.weight", "acc_bottleneck.0.bias",
    "cls_ln.weight", "cls_ln.bias", "delta_ln.weight", "delta_ln.bias", "acc_ln.weight", "acc_ln.bias",
    "lin0.0.weight", "lin0.0.bias", "attention_head.weight", "attention_head.bias",
    "lin1.weight", "lin1.bias", "lin2.weight", "lin2.bias",
    "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
    "lstm.weight_ih_l0_reverse", "lstm.weight_hh_l0_reverse", "lstm.bias_ih_l0_reverse", "lstm.bias_hh_l0_reverse",
)


def make_head_state(in_features=768, out_features=9, bottleneck=128, lstm_hidden=64, seed=0,
                    scale=1.0, lstm_layers=1, use_acceleration=True) -> Dict[str, torch.Tensor]:
    """Deterministic head weights in the reference state_dict layout (workthreads.py:856 model.pth keys),
    drawn from numpy's default_rng so the fixture can be regenerated anywhere without torch's RNG stream.
    Uniform(+-scale/sqrt(fan_in)) like nn.Linear/nn.LSTM defaults; LayerNorm gains jittered around 1."""
    rng = np.random.default_rng(seed)

    def u(shape, fan_in):
        b = scale / math.sqrt(fan_in)
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    Hs, Bn, Fi, Co = lstm_hidden, bottleneck, in_features, out_features
    sd = {"gate": torch.tensor(0.2), "attention_temp": torch.tensor(1.0)}
    streams = ("cls", "delta", "acc") if use_acceleration else ("cls", "delta")
    for s in streams:
        sd[f"{s}_bottleneck.0.weight"] = u((Bn, Fi), Fi)
        sd[f"{s}_bottleneck.0.bias"] = u((Bn,), Fi)
        sd[f"{s}_ln.weight"] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(Bn)).astype(np.float32))
        sd[f"{s}_ln.bias"] = torch.from_numpy((0.1 * rng.standard_normal(Bn)).astype(np.float32))
    aug = len(streams) * Bn
    sd["lin0.0.weight"], sd["lin0.0.bias"] = u((256, aug), aug), u((256,), aug)
    sd["attention_head.weight"], sd["attention_head.bias"] = u((1, 2 * Hs), 2 * Hs), u((1,), 2 * Hs)
    sd["lin1.weight"], sd["lin1.bias"] = u((Co, Fi), Fi), u((Co,), Fi)
    sd["lin2.weight"], sd["lin2.bias"] = u((Co, 2 * Hs), 2 * Hs), u((Co,), 2 * Hs)
    for layer in range(lstm_layers):  # nn.LSTM: layer k > 0 reads the [fwd | rev] outputs of layer k-1
        kin = 256 if layer == 0 else 2 * Hs
        for sfx in ("", "_reverse"):
            sd[f"lstm.weight_ih_l{layer}{sfx}"] = u((4 * Hs, kin), Hs)
            sd[f"lstm.weight_hh_l{layer}{sfx}"] = u((4 * Hs, Hs), Hs)
            sd[f"lstm.bias_ih_l{layer}{sfx}"] = u((4 * Hs,), Hs)
            sd[f"lstm.bias_hh_l{layer}{sfx}"] = u((4 * Hs,), Hs)
    return sd


def robust_deltas(x: torch.Tensor, alpha: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """classifier_head.py:102-117: window-local EMA, then first/second differences of the sequence
    reflect-padded by 2 on the left (replicate when T < 3)."""
    B, T, Cc = x.shape
    s = torch.zeros_like(x)
    s[:, 0] = x[:, 0]
    for t in range(1, T):
        s[:, t] = torch.lerp(s[:, t - 1], x[:, t], alpha)
    mode = "reflect" if T >= 3 else "replicate"
    padded = F.pad(s.permute(0, 2, 1), (2, 0), mode).permute(0, 2, 1)
    dx = padded[:, 1:] - padded[:, :-1]
    ddx = dx[:, 1:] - dx[:, :-1]
    return s, dx[:, 1:], ddx


def _lstm_direction(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of nn.LSTM (gate order i, f, g, o; zero initial state), batch-first, all time steps."""
    B, T, _ = x.shape
    Hs = w_hh.shape[1]
    h = x.new_zeros(B, Hs)
    c = x.new_zeros(B, Hs)
    out = x.new_zeros(B, T, Hs)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = x[:, t] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f, gg, o = g.split(Hs, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def head_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, seq_len: int = 31, center_window: int = 5,
                 ema_alpha: float = 0.3, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """ClassifierLSTMDeltas.forward in eval mode (classifier_head.py:150-172): x [B,T,F] -> (logits [B,C],
    rawm [B,2Hs]).  The number of LSTM layers, the hidden size and use_acceleration are read off the state dict
    (as the reference's bundle loader does, workthreads.py:416-425)."""
    p = {k: v.to(dtype) for k, v in sd.items()}
    x = x.to(dtype)
    hsl, sw = seq_len // 2, center_window
    cls_s, d_s, a_s = robust_deltas(x, ema_alpha)
    L = x.shape[1]
    l, r = max(0, hsl - sw), min(L, hsl + sw + 1)

    # forward_linear (classifier_head.py:119-129)
    if l >= r:
        idx = min(max(0, L // 2), L - 1)
        linear_logits = cls_s[:, idx] @ p["lin1.weight"].T + p["lin1.bias"]
    else:
        linear_logits = (cls_s[:, l:r] @ p["lin1.weight"].T + p["lin1.bias"]).mean(dim=1)

    def bott(stream, name):
        y = F.gelu(stream @ p[f"{name}_bottleneck.0.weight"].T + p[f"{name}_bottleneck.0.bias"])
        return F.layer_norm(y, (y.shape[-1],), p[f"{name}_ln.weight"], p[f"{name}_ln.bias"], 1e-5)

    parts = [bott(cls_s, "cls"), bott(d_s, "delta")]
    if "acc_bottleneck.0.weight" in p:  # use_acceleration (classifier_head.py:163-167)
        parts.append(bott(a_s, "acc"))
    aug = torch.cat(parts, dim=-1)
    z = F.gelu(aug @ p["lin0.0.weight"].T + p["lin0.0.bias"])
    z = z - z.mean(dim=1, keepdim=True)

    out, layer = z, 0
    while f"lstm.weight_ih_l{layer}" in p:  # stacked bidirectional layers (nn.LSTM num_layers; no dropout in eval)
        k = f"_l{layer}"
        fwd = _lstm_direction(out, p["lstm.weight_ih" + k], p["lstm.weight_hh" + k], p["lstm.bias_ih" + k],
                              p["lstm.bias_hh" + k], False)
        bwd = _lstm_direction(out, p["lstm.weight_ih" + k + "_reverse"], p["lstm.weight_hh" + k + "_reverse"],
                              p["lstm.bias_ih" + k + "_reverse"], p["lstm.bias_hh" + k + "_reverse"], True)
        out = torch.cat([fwd, bwd], dim=-1)
        layer += 1

    # forward_lstm (classifier_head.py:131-148)
    if l >= r:
        idx = min(max(0, L // 2), L - 1)
        rawm = out[:, idx]
    else:
        win = out[:, l:r]
        temp = F.softplus(p["attention_temp"]) + 1e-3
        scores = (win @ p["attention_head.weight"].T + p["attention_head.bias"]).squeeze(-1) / temp
        rawm = (torch.softmax(scores, dim=1).unsqueeze(-1) * win).sum(dim=1)
    lstm_logits = rawm @ p["lin2.weight"].T + p["lin2.bias"]
    final = torch.lerp(linear_logits, lstm_logits, torch.sigmoid(p["gate"]))
    return final, rawm


def infer_windows(emb: np.ndarray, sd: Dict[str, torch.Tensor], seq_len: int = 31, temperature: float = 1.0,
                  batch: int = 512, chunk: int = 20000, return_logits: bool = False, dtype=torch.float32,
                  **head_kw):
    """infer_file's numeric core (cbas.py:497-551) over an in-memory `cls` array [N,F] (float16 as stored):
    chunks of `chunk` frames read with +-seq_len//2 context, replicate padding at the video ends, stride-1
    windows, batches of `batch`, probs = softmax(logits / max(1e-3, T)).  Returns probs [N,C] float32."""
    total = emb.shape[0]
    half = seq_len // 2
    probs_all, logits_all = [], []
    for start in range(0, total, chunk):
        end = min(start + chunk, total)
        rs, re_ = max(0, start - half), min(total, end + half)
        ct = torch.from_numpy(np.asarray(emb[rs:re_])).float()
        if start < half:
            pad = half - start
            if pad > 0:
                ct = torch.cat([ct[0:1].repeat(pad, 1), ct], dim=0)
        if end > total - half:
            pad = half - (total - end)
            if pad > 0:
                ct = torch.cat([ct, ct[-1:].repeat(pad, 1)], dim=0)
        n_t = end - start
        for b0 in range(0, n_t, batch):
            b1 = min(b0 + batch, n_t)
            win = torch.stack([ct[i:i + seq_len] for i in range(b0, b1)])
            with torch.no_grad():
                logits, _ = head_forward(sd, win, seq_len=seq_len, dtype=dtype, **head_kw)
                pr = torch.softmax(logits / max(1e-3, temperature), dim=1)
            probs_all.append(pr.float().numpy())
            logits_all.append(logits.float().numpy())
    probs = np.concatenate(probs_all) if probs_all else np.zeros((0, 0), np.float32)
    if return_logits:
        return probs, np.concatenate(logits_all)
    return probs
