"""ClassifierLSTMDeltas - host-side mirror of the reference head over the sm_100a C ABI.

Mirrors `backend/classifier_head.py:57-172`: same constructor signature, same parameter names and shapes (a
reference `model.pth` loads with `load_state_dict`, strict or not, exactly as `workthreads.py:441` does), same
`forward(x[B,T,F]) -> (final_logits[B,C], rawm[B,2*Hs])` in eval mode.  The arithmetic runs in libcbas_b200.so
(csrc/head.cu).  `infer_embeddings` is the fast path `infer_file` uses: the whole `cls` array in, one
probability row per frame out, with infer_file's windowing, edge padding and temperature softmax on the GPU.

Eval mode is the product path (the native kernels, no autograd).  In train mode (`model.train()`, used only by
`cbas_b200.training.train_lstm_model`) forward() is a differentiable restatement of the same function
in torch ops on the GPU - dropout, autograd and the cuDNN LSTM are library code there, on purpose: the gradient step
is not on the streamed encode / inference path this library accelerates, while every evaluation pass of a training
run and the model it returns go through the native kernels again.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class ClassifierLSTMDeltas(nn.Module):
    def __init__(self, in_features, out_features, seq_len=31, bottleneck_dim=128,
                 dropout_p=0.15, use_acceleration=True, ema_alpha=0.3, center_window_size=5,
                 lstm_hidden_size=64, lstm_layers=1):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.seq_len, self.sw = seq_len, center_window_size
        self.hsl = seq_len // 2
        self.use_acceleration = use_acceleration
        self.ema_alpha = ema_alpha
        self.bottleneck_dim = bottleneck_dim
        self.lstm_hidden_size, self.lstm_layers = lstm_hidden_size, lstm_layers

        # identical module tree to the reference so state_dict keys match (workthreads.py:856 model.pth)
        self.cls_bottleneck = nn.Sequential(nn.Linear(in_features, bottleneck_dim), nn.GELU(), nn.Dropout(0.1))
        self.delta_bottleneck = nn.Sequential(nn.Linear(in_features, bottleneck_dim), nn.GELU(), nn.Dropout(0.1))
        if use_acceleration:
            self.acc_bottleneck = nn.Sequential(nn.Linear(in_features, bottleneck_dim), nn.GELU(), nn.Dropout(0.1))
        self.cls_ln = nn.LayerNorm(bottleneck_dim)
        self.delta_ln = nn.LayerNorm(bottleneck_dim)
        if use_acceleration:
            self.acc_ln = nn.LayerNorm(bottleneck_dim)
        augmented = bottleneck_dim * 3 if use_acceleration else bottleneck_dim * 2
        self.lin0 = nn.Sequential(nn.Linear(augmented, 256), nn.GELU(), nn.Dropout(dropout_p))
        self.gate = nn.Parameter(torch.tensor(0.2))
        self.attention_head = nn.Linear(lstm_hidden_size * 2, 1)
        self.attention_temp = nn.Parameter(torch.tensor(1.0))
        self.lin1 = nn.Linear(in_features, out_features)
        self.lin2 = nn.Linear(lstm_hidden_size * 2, out_features)
        self.lstm = nn.LSTM(256, lstm_hidden_size, num_layers=lstm_layers, batch_first=True, bidirectional=True)
        self._native = None
        self._native_key = None
        self._stream_mats = None
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def train(self, mode: bool = True):
        """Training mode switches forward() to the differentiable torch path and makes the parameters trainable;
        eval() goes back to the native kernels (parameters frozen, as a loaded model.pth is used)."""
        super().train(mode)
        for p in self.parameters():
            p.requires_grad_(bool(mode))
        return self

    # ---- differentiable path (training only) --------------------------------------------------------------
    def _time_operators(self, T: int, device, dtype):
        """The three temporal streams of classifier_head.py:100-116 as constant [T,T] operators on the time axis:
        S = EMA smoothing (x_s[0] = x[0], x_s[t] = lerp(x_s[t-1], x[t], alpha)), D1 = first difference of the
        smoothed stream with the two-frame reflected head (delta[0] = x_s[0] - x_s[1]), D2 = second difference
        (acc[t] = d[t+1] - d[t] over the padded differences)."""
        key = (T, str(device), dtype)
        if self._stream_mats is not None and self._stream_mats[0] == key:
            return self._stream_mats[1]
        a = float(self.ema_alpha)
        S = torch.zeros(T, T, dtype=torch.float64)
        S[0, 0] = 1.0
        for t in range(1, T):
            S[t] = (1.0 - a) * S[t - 1]
            S[t, t] += a
        # padded sequence p = [x2, x1, x0, x1, ..., x_{T-1}] (reflect) or [x0, x0, x0, x1, ...] (replicate, T < 3)
        P = torch.zeros(T + 2, T, dtype=torch.float64)
        head = (2, 1) if T >= 3 else (0, 0)
        P[0, head[0]] = 1.0
        P[1, head[1]] = 1.0
        P[2:] = torch.eye(T, dtype=torch.float64)
        dP = P[1:] - P[:-1]            # [T+1, T]: all padded first differences
        D1 = dP[1:]                     # delta stream: the last T of them
        D2 = dP[1:] - dP[:-1]           # acceleration stream
        mats = tuple((m @ S if i else S).to(device=device, dtype=dtype) for i, m in enumerate((S, D1, D2)))
        self._stream_mats = (key, mats)
        return mats

    def _forward_autograd(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        B, T, _ = x.shape
        x = x.float()
        S, D1, D2 = self._time_operators(T, x.device, x.dtype)
        smooth = torch.einsum("ts,bsf->btf", S, x)
        delta = torch.einsum("ts,bsf->btf", D1, x)
        lo, hi = max(0, self.hsl - self.sw), min(T, self.hsl + self.sw + 1)
        centre = slice(lo, hi) if lo < hi else slice(min(max(0, T // 2), T - 1), min(max(0, T // 2), T - 1) + 1)
        # linear branch: mean over the centre window of lin1(smoothed frames) (classifier_head.py:118-128)
        linear_logits = self.lin1(smooth[:, centre]).mean(dim=1)
        parts = [self.cls_ln(self.cls_bottleneck(smooth)), self.delta_ln(self.delta_bottleneck(delta))]
        if self.use_acceleration:
            parts.append(self.acc_ln(self.acc_bottleneck(torch.einsum("ts,bsf->btf", D2, x))))
        z = self.lin0(torch.cat(parts, dim=-1))
        z = z - z.mean(dim=1, keepdim=True)
        out, _ = self.lstm(z)
        win = out[:, centre]
        if lo < hi:
            temp = torch.nn.functional.softplus(self.attention_temp) + 1e-3
            w = torch.softmax(self.attention_head(win).squeeze(-1) / temp, dim=1).unsqueeze(-1)
            rawm = (w * win).sum(dim=1)
        else:
            rawm = win[:, 0]
        lstm_logits = self.lin2(rawm)
        return torch.lerp(linear_logits, lstm_logits, torch.sigmoid(self.gate)), rawm

    # ---- native handle management -----------------------------------------------------------------------
    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .float() move the parameters: rebuild the handle lazily
        self._drop_native()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._drop_native()
        return super().load_state_dict(*a, **k)

    def _drop_native(self):
        h = getattr(self, "_native", None)
        if h:
            try:
                _lib.lib().cbas_b200_head_destroy(h)
            except Exception:
                pass
        self._native = None
        self._native_key = None

    def __del__(self):
        try:
            self._drop_native()
        except Exception:
            pass

    def _get_native(self) -> C.c_void_p:
        dev = self.lin1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("cbas_b200.ClassifierLSTMDeltas runs on CUDA (sm_100a) only; move it with .to('cuda')")
        key = (dev, tuple(int(p._version) for p in self.parameters()))
        if self._native is not None and self._native_key == key:
            return self._native
        self._drop_native()
        if self.lstm_hidden_size not in (64, 128) or self.lstm_layers not in (1, 2):
            raise NotImplementedError("cbas_b200 head: lstm_hidden_size must be 64 or 128 and lstm_layers 1 or 2")
        lib = _lib.lib()
        cfg = _lib.HeadCfg(self.in_features, self.out_features, self.seq_len, self.bottleneck_dim,
                           self.lstm_hidden_size, self.sw, float(self.ema_alpha), int(bool(self.use_acceleration)),
                           int(self.lstm_layers))
        keep = []

        def ptr(t: torch.Tensor) -> int:
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.HeadWeights()
        w.cls_w, w.cls_b = ptr(self.cls_bottleneck[0].weight), ptr(self.cls_bottleneck[0].bias)
        w.delta_w, w.delta_b = ptr(self.delta_bottleneck[0].weight), ptr(self.delta_bottleneck[0].bias)
        if self.use_acceleration:
            w.acc_w, w.acc_b = ptr(self.acc_bottleneck[0].weight), ptr(self.acc_bottleneck[0].bias)
            w.acc_ln_g, w.acc_ln_b = ptr(self.acc_ln.weight), ptr(self.acc_ln.bias)
        w.cls_ln_g, w.cls_ln_b = ptr(self.cls_ln.weight), ptr(self.cls_ln.bias)
        w.delta_ln_g, w.delta_ln_b = ptr(self.delta_ln.weight), ptr(self.delta_ln.bias)
        w.lin0_w, w.lin0_b = ptr(self.lin0[0].weight), ptr(self.lin0[0].bias)
        w.lin1_w, w.lin1_b = ptr(self.lin1.weight), ptr(self.lin1.bias)
        w.lin2_w, w.lin2_b = ptr(self.lin2.weight), ptr(self.lin2.bias)
        w.att_w, w.att_b = ptr(self.attention_head.weight), ptr(self.attention_head.bias)
        w.w_ih_f, w.w_hh_f = ptr(self.lstm.weight_ih_l0), ptr(self.lstm.weight_hh_l0)
        w.b_ih_f, w.b_hh_f = ptr(self.lstm.bias_ih_l0), ptr(self.lstm.bias_hh_l0)
        w.w_ih_r, w.w_hh_r = ptr(self.lstm.weight_ih_l0_reverse), ptr(self.lstm.weight_hh_l0_reverse)
        w.b_ih_r, w.b_hh_r = ptr(self.lstm.bias_ih_l0_reverse), ptr(self.lstm.bias_hh_l0_reverse)
        if self.lstm_layers == 2:
            w.w_ih_f1, w.w_hh_f1 = ptr(self.lstm.weight_ih_l1), ptr(self.lstm.weight_hh_l1)
            w.b_ih_f1, w.b_hh_f1 = ptr(self.lstm.bias_ih_l1), ptr(self.lstm.bias_hh_l1)
            w.w_ih_r1, w.w_hh_r1 = ptr(self.lstm.weight_ih_l1_reverse), ptr(self.lstm.weight_hh_l1_reverse)
            w.b_ih_r1, w.b_hh_r1 = ptr(self.lstm.bias_ih_l1_reverse), ptr(self.lstm.bias_hh_l1_reverse)
        w.gate = float(self.gate.detach().cpu())
        w.attention_temp = float(self.attention_temp.detach().cpu())
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            torch.cuda.synchronize(dev)  # the staged weight copies above must have landed before create() reads them
            _lib.check(lib.cbas_b200_head_create(C.byref(cfg), C.byref(w), C.byref(handle)), "head_create")
            torch.cuda.synchronize(dev)  # create() copies the weights; `keep` may go after this
        self._native, self._native_key = handle, key
        return handle

    # ---- reference-compatible call ------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x [B,T,F] -> (final_logits [B,C], rawm [B,2*Hs]).  Eval mode: the native kernels; train mode: the
        differentiable torch restatement (see the module docstring)."""
        B, T, Fdim = x.shape
        if T != self.seq_len or Fdim != self.in_features:
            raise ValueError(f"expected windows of [{self.seq_len}, {self.in_features}], got [{T}, {Fdim}]")
        if self.training:
            if self.lin1.weight.device.type != "cuda":
                raise RuntimeError("cbas_b200.ClassifierLSTMDeltas runs on CUDA (sm_100a) only; move it with .to('cuda')")
            return self._forward_autograd(x.to(self.lin1.weight.device))
        h = self._get_native()
        dev = self.lin1.weight.device
        xd = x.to(device=dev, dtype=torch.float32).contiguous()
        logits = torch.empty(B, self.out_features, device=dev, dtype=torch.float32)
        rawm = torch.empty(B, 2 * self.lstm_hidden_size, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().cbas_b200_head_forward_windows(
            h, xd.data_ptr(), B, logits.data_ptr(), rawm.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
            "head_forward_windows")
        return logits, rawm

    # ---- fast path ----------------------------------------------------------------------------------------
    def infer_embeddings(self, emb: torch.Tensor, temperature: float = 1.0,
                         return_logits: bool = False):
        """emb: float16 [N,F] on the head's device (the `cls` dataset as stored).  Returns probs float32 [N,C]
        (and final logits when asked): for every frame the window [f-T/2, f+T/2] with replicate padding at both
        ends, softmax(logits / max(1e-3, temperature)) - infer_file's numeric core (cbas.py:497-551)."""
        dev = self.lin1.weight.device
        if emb.dtype != torch.float16 or emb.dim() != 2 or emb.shape[1] != self.in_features:
            raise ValueError(f"infer_embeddings expects float16 [N,{self.in_features}]")
        emb = emb.to(dev).contiguous()
        n = emb.shape[0]
        probs = torch.empty(n, self.out_features, device=dev, dtype=torch.float32)
        logits = torch.empty(n, self.out_features, device=dev, dtype=torch.float32) if return_logits else None
        if n:
            h = self._get_native()
            _lib.check(_lib.lib().cbas_b200_head_infer(
                h, emb.data_ptr(), n, float(temperature), probs.data_ptr(),
                logits.data_ptr() if logits is not None else None, torch.cuda.current_stream(dev).cuda_stream),
                "head_infer")
        return (probs, logits) if return_logits else probs


def synthetic_head_state_dict(in_features: int = 768, out_features: int = 9, bottleneck_dim: int = 128,
                              lstm_hidden_size: int = 64, lstm_layers: int = 1, use_acceleration: bool = True,
                              seed: int = 0, scale: float = 2.0):
    """Random head weights in the reference's `model.pth` layout (workthreads.py:856): nn.Linear / nn.LSTM-style
    uniform(+-scale / sqrt(fan_in)), LayerNorm gains near 1.  For benchmarks and smoke runs (no trained bundle ships
    with the repo); load with `ClassifierLSTMDeltas(...).load_state_dict(sd)`."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan_in):
        b = scale / fan_in ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    Hs, Bn, Fi, Co = lstm_hidden_size, bottleneck_dim, in_features, out_features
    sd = {"gate": torch.tensor(0.2), "attention_temp": torch.tensor(1.0)}
    streams = ("cls", "delta", "acc") if use_acceleration else ("cls", "delta")
    for s in streams:
        sd[f"{s}_bottleneck.0.weight"], sd[f"{s}_bottleneck.0.bias"] = u((Bn, Fi), Fi), u((Bn,), Fi)
        sd[f"{s}_ln.weight"] = 1.0 + 0.1 * torch.randn(Bn, generator=g)
        sd[f"{s}_ln.bias"] = 0.1 * torch.randn(Bn, generator=g)
    aug = len(streams) * Bn
    sd["lin0.0.weight"], sd["lin0.0.bias"] = u((256, aug), aug), u((256,), aug)
    sd["attention_head.weight"], sd["attention_head.bias"] = u((1, 2 * Hs), 2 * Hs), u((1,), 2 * Hs)
    sd["lin1.weight"], sd["lin1.bias"] = u((Co, Fi), Fi), u((Co,), Fi)
    sd["lin2.weight"], sd["lin2.bias"] = u((Co, 2 * Hs), 2 * Hs), u((Co,), 2 * Hs)
    for layer in range(lstm_layers):
        kin = 256 if layer == 0 else 2 * Hs
        for sfx in ("", "_reverse"):
            sd[f"lstm.weight_ih_l{layer}{sfx}"], sd[f"lstm.weight_hh_l{layer}{sfx}"] = u((4 * Hs, kin), Hs), u((4 * Hs, Hs), Hs)
            sd[f"lstm.bias_ih_l{layer}{sfx}"], sd[f"lstm.bias_hh_l{layer}{sfx}"] = u((4 * Hs,), Hs), u((4 * Hs,), Hs)
    return sd


def actogram_bins(probs: torch.Tensor, behavior: int, threshold: float, bin_frames: int) -> torch.Tensor:
    """Numeric core of Actogram.__init__ (cbas.py:969-999) on the GPU: probs [N,C] (CUDA; float32 for probabilities
    that stayed on the device, float64 for tables parsed from CSV files, compared in float64 like the reference) ->
    int32 bin counts [ceil(N / bin_frames)]."""
    if probs.dtype not in (torch.float32, torch.float64) or probs.dim() != 2 or not probs.is_cuda:
        raise ValueError("actogram_bins expects a float32 or float64 [N,C] CUDA tensor")
    probs = probs.contiguous()
    n, c = probs.shape
    nb = (n + bin_frames - 1) // bin_frames if bin_frames > 0 else 0
    bins = torch.zeros(nb, device=probs.device, dtype=torch.int32)
    if n and nb:
        fn = _lib.lib().cbas_b200_actogram_bins if probs.dtype == torch.float32 else _lib.lib().cbas_b200_actogram_bins_f64
        _lib.check(fn(probs.data_ptr(), n, c, int(behavior), float(threshold), int(bin_frames), bins.data_ptr(),
                      torch.cuda.current_stream(probs.device).cuda_stream), "actogram_bins")
    return bins
