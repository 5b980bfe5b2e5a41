"""Head training behind the reference's surface: `train_lstm_model` (backend/cbas.py:1274-1422) and
`fit_temperature` (backend/workthreads.py:103-137).

What runs where.  The gradient step - forward with dropout, cross entropy + the decorrelation penalty on the pooled
LSTM state (cbas.py:1338-1342), backward, Adam - goes through `ClassifierLSTMDeltas` in train mode, i.e. torch autograd
and the cuDNN LSTM on the GPU: library code, deliberately (training is not on the streamed encode / inference path
this library accelerates; SURVEY.md ranks it last).  Everything else a training run does with the head is inference
and runs on the native kernels: the per-epoch prediction passes over the training and validation sets (most of an
epoch's forward passes), the logits `fit_temperature` calibrates on, and the model that is returned (eval mode).
There is no CPU path: a non-CUDA device raises.

Same arguments, return values and early-stopping rule as the reference.  One deliberate difference: the reference keeps
`model.state_dict().copy()` as the best state, a shallow copy whose tensors keep training, so it returns the LAST
epoch's weights under the best epoch's number; here the best epoch's weights are cloned and returned.
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .classifier_head import ClassifierLSTMDeltas


class PerformanceReport:
    """Per-epoch sklearn reports and confusion matrices (cbas.py:1267-1272)."""

    def __init__(self, train_report: dict, train_cm: np.ndarray, val_report: dict, val_cm: np.ndarray):
        self.train_report, self.train_cm = train_report, train_cm
        self.val_report, self.val_cm = val_report, val_cm


def collate_fn(batch):
    """Drop samples whose label is -1 (failed loads) and stack the rest (cbas.py:1253-1260)."""
    keep = [(d, l) for d, l in batch if int(l) != -1]
    if not keep:
        return torch.tensor([]), torch.tensor([])
    return torch.stack([d for d, _ in keep]), torch.stack([torch.as_tensor(l) for _, l in keep])


def decorrelation_penalty(rawm: torch.Tensor) -> torch.Tensor:
    """Sum of squared off-diagonal entries of the batch covariance of the pooled LSTM state (cbas.py:1262-1265,
    1338-1342); zero for a batch of one."""
    if rawm.ndim != 2 or rawm.shape[0] < 2:
        return rawm.new_zeros(())
    c = rawm - rawm.mean(dim=0)
    cov = (c.T @ c) / (rawm.shape[0] - 1)
    return (cov.pow(2).sum() - cov.diagonal().pow(2).sum())


def _log(msg: str) -> None:
    try:
        from .workthreads import log_message
        log_message(msg, "INFO")
    except Exception:
        print(msg)


def _cuda_device(device) -> torch.device:
    device = torch.device(device) if device is not None else torch.device("cuda")
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("cbas_b200.training runs on CUDA only (no CPU fallback)")
    return device


@torch.no_grad()
def _predict(model: ClassifierLSTMDeltas, loader, device, cancel_event=None) -> Tuple[List[int], List[int]]:
    """Labels and argmax predictions over a loader, native kernels (model in eval mode)."""
    actual, predicted = [], []
    for d, l in loader:
        if cancel_event is not None and cancel_event.is_set():
            break
        if d.numel() == 0:
            continue
        logits, _ = model(d.to(device).float())
        actual.extend(np.asarray(l.cpu()).tolist())
        predicted.extend(logits.argmax(1).cpu().numpy().tolist())
    return actual, predicted


def train_lstm_model(train_set, test_set, seq_len: int, behaviors: list, cancel_event: threading.Event,
                     batch_size=512, lr=1e-4, epochs=10, device=None, class_weights=None, patience=3,
                     progress_callback: Optional[Callable[[str], None]] = None, optimization_target="weighted avg",
                     weight_decay=0.0, label_smoothing=0.0, lstm_hidden_size=64, lstm_layers=1,
                     in_features: Optional[int] = None) -> tuple:
    """Train a ClassifierLSTMDeltas head on (window[T,F], label) samples.  Returns (model in eval mode, per-epoch
    PerformanceReports, best epoch); (None, reports, best epoch) when cancelled; (None, None, -1) for an empty training
    set or when nothing was learnt.  `in_features` defaults to the width of the first training sample (the reference
    hard-codes 768)."""
    from sklearn.metrics import classification_report, confusion_matrix

    if len(train_set) == 0:
        return None, None, -1
    device = _cuda_device(device)
    if in_features is None:
        in_features = int(train_set[0][0].shape[-1])
    mk = dict(collate_fn=collate_fn, num_workers=0)
    train_loader = torch.utils.data.DataLoader(train_set, batch_size, shuffle=True, pin_memory=True, drop_last=False, **mk)
    test_loader = torch.utils.data.DataLoader(test_set, batch_size, shuffle=False, **mk) if test_set and len(test_set) > 0 else None

    def new_model():
        return ClassifierLSTMDeltas(in_features=in_features, out_features=len(behaviors), seq_len=seq_len,
                                    lstm_hidden_size=lstm_hidden_size, lstm_layers=lstm_layers)

    model = new_model().to(device)
    _log(f"Successfully instantiated model architecture: {type(model).__name__}")
    _log(f"Training hyperparameters: lr {lr}, weight decay {weight_decay}, label smoothing {label_smoothing}, "
         f"LSTM hidden size {lstm_hidden_size}, LSTM layers {lstm_layers}")
    model.train()  # parameters become trainable before the optimizer sees them
    optimizer = torch.optim.Adam([
        {"params": [p for name, p in model.named_parameters() if name != "gate"]},
        {"params": [model.gate], "weight_decay": 1e-3},
    ], lr=lr, weight_decay=weight_decay)
    weights = torch.tensor(class_weights, dtype=torch.float, device=device) if class_weights is not None else None
    criterion = nn.CrossEntropyLoss(weight=weights, label_smoothing=label_smoothing)
    labels = range(len(behaviors))

    best_f1, best_state, best_epoch = -1.0, None, -1
    reports: List[PerformanceReport] = []
    stale = 0
    for e in range(epochs):
        if cancel_event.is_set():
            return None, reports, best_epoch
        if progress_callback:
            progress_callback(f"Training Epoch {e + 1}/{epochs}...")
        model.train()
        for i, (d, l) in enumerate(train_loader):
            if cancel_event.is_set():
                break
            if d.numel() == 0:
                continue
            d, l = d.to(device).float(), l.to(device)
            optimizer.zero_grad()
            logits, rawm = model(d)
            loss = criterion(logits, l) + decorrelation_penalty(rawm)
            loss.backward()
            optimizer.step()
            if i % 50 == 0:
                print(f"[Epoch {e + 1}/{epochs} Batch {i}/{len(train_loader)}] Loss: {loss.item():.4f}")

        model.eval()  # prediction passes: native kernels
        actual, predicted = _predict(model, train_loader, device)
        if not actual:
            stale += 1
            if stale >= patience:
                break
            continue
        kw = dict(target_names=behaviors, output_dict=True, zero_division=0, labels=labels)
        train_report = classification_report(actual, predicted, **kw)
        train_cm = confusion_matrix(actual, predicted, labels=labels)
        val_report, val_cm = {}, np.array([])
        if test_loader:
            v_actual, v_predicted = _predict(model, test_loader, device, cancel_event)
            if v_actual:
                val_report = classification_report(v_actual, v_predicted, **kw)
                val_cm = confusion_matrix(v_actual, v_predicted, labels=labels)
        reports.append(PerformanceReport(train_report, train_cm, val_report, val_cm))

        val_f1 = val_report.get(optimization_target, {}).get("f1-score", -1.0)
        train_f1 = train_report.get(optimization_target, {}).get("f1-score", -1.0)
        shown = f"{val_f1:.4f}" if test_loader else "N/A"
        if progress_callback:
            progress_callback(f"Epoch {e + 1} Val F1: {shown}")
        print(f"--- Epoch {e + 1} | Train F1: {train_f1:.4f} | Val F1: {shown} ({optimization_target}) ---")
        if val_f1 > best_f1:
            best_f1, best_epoch = val_f1, e
            best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
            stale = 0
        else:
            stale += 1
        if test_loader and stale >= patience:
            _log(f"Early stopping triggered at epoch {e + 1}.")
            break

    if best_state is None and epochs > 0 and not test_loader:
        best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        best_epoch = epochs - 1
    if best_state:
        final = new_model()
        final.load_state_dict(best_state)
        return final.eval(), reports, best_epoch
    return None, None, -1


def fit_temperature(model, val_loader, device) -> float:
    """Temperature that calibrates the head's confidence on a validation loader (workthreads.py:103-137): L-BFGS on
    softplus(T) + 1e-3 (clamped at 10) minimising the cross entropy of logits / temperature; 1.0 for an empty loader.
    The logits come from the native kernels; the scalar fit is a few dozen torch ops."""
    device = _cuda_device(device)
    model.to(device)
    model.eval()
    chunks, labels = [], []
    with torch.no_grad():
        for d, l in val_loader:
            if d.numel() == 0:
                continue
            logits, _ = model(d.to(device).float())
            chunks.append(logits)
            labels.append(torch.as_tensor(l))
    if not chunks:
        return 1.0
    logits = torch.cat(chunks).detach()
    target = torch.cat(labels).to(device)
    T = torch.nn.Parameter(torch.ones(1, device=device))
    opt = torch.optim.LBFGS([T], lr=0.01, max_iter=50)
    ce = nn.CrossEntropyLoss()

    def temperature():
        return torch.clamp(torch.nn.functional.softplus(T) + 1e-3, max=10.0)

    def closure():
        opt.zero_grad()
        loss = ce(logits / temperature(), target)
        loss.backward()
        return loss

    opt.step(closure)
    return float(temperature().detach().cpu().item())
