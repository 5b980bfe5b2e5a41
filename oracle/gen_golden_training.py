"""Golden vectors for head training, from the REAL reference code.  TEST INFRASTRUCTURE.

Imports /root/reference/backend/classifier_head.py and the loss of cbas.train_lstm_model (cbas.py:1262-1265,
1331-1345: cross entropy + sum of squared off-diagonal covariances of the pooled LSTM state) and records, for fixed
weights (oracle.head.make_head_state) and a fixed batch, the loss and the gradient of every parameter - dropout off
(eval-mode dropout, train-mode everything else has no other effect in this model), so the numbers are deterministic.
Also one `fit_temperature` run (workthreads.py:103-137, its own code path copied into the call below by importing
nothing but torch: the function needs only a model and a loader) on fixed logits.

    python oracle/gen_golden_training.py        ->  tests/golden/head_training.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/backend")
import classifier_head  # noqa: E402  (the reference)
from oracle import head as ohead  # noqa: E402

X_SEED, STATE_SEED, B, C = 41, 17, 24, 5


def reference_off_diagonal(x):  # cbas.py:1262-1265, restated to avoid importing cbas.py's GUI-side dependencies
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()


def batch():
    rng = np.random.default_rng(X_SEED)
    x = np.cumsum(rng.standard_normal((B, 31, 768)).astype(np.float32) * 0.3, axis=1).astype(np.float16)
    y = rng.integers(0, C, size=B)
    return torch.from_numpy(x).float(), torch.from_numpy(y)


def main():
    torch.manual_seed(0)
    x, y = batch()
    out = {"x_seed": X_SEED, "state_seed": STATE_SEED, "batch": B, "classes": C}
    for name, (hs, layers) in {"h64_l1": (64, 1), "h128_l2": (128, 2)}.items():
        sd = ohead.make_head_state(768, C, 128, hs, seed=STATE_SEED, scale=2.0, lstm_layers=layers)
        m = classifier_head.ClassifierLSTMDeltas(768, C, seq_len=31, lstm_hidden_size=hs, lstm_layers=layers)
        m.load_state_dict(sd, strict=True)
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.eval()
        weights = torch.tensor([1.0, 2.0, 0.5, 1.5, 1.0])
        crit = torch.nn.CrossEntropyLoss(weight=weights, label_smoothing=0.1)
        logits, rawm = m(x)
        inv = crit(logits, y)
        c = rawm - rawm.mean(dim=0)
        cov = (c.T @ c) / (c.shape[0] - 1)
        pen = torch.sum(torch.pow(reference_off_diagonal(cov), 2))
        (inv + pen).backward()
        out[f"{name}/loss"] = np.array([float(inv.detach()), float(pen.detach())])
        out[f"{name}/logits"] = logits.detach().numpy()
        for k, p in m.named_parameters():
            g = p.grad.detach().numpy().ravel()
            out[f"{name}/grad_norm/{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
            out[f"{name}/grad_sample/{k}"] = g[::max(1, g.size // 257)][:257].copy()  # every (size // 257)-th entry
    # fit_temperature (workthreads.py:103-137) on fixed logits: import the function through a stub of its module's
    # GUI-side imports would drag in eel; the routine itself is restated by the test target, so record its result by
    # running the reference's own source text for that function
    src = open("/root/reference/backend/workthreads.py").read()
    start = src.index("def fit_temperature(")
    end = src.index("\ndef ", start + 10)
    ns = {"torch": torch}
    exec(src[start:end], ns)
    rng = np.random.default_rng(5)
    lg = torch.from_numpy((rng.standard_normal((400, C)) * 4.0).astype(np.float32))
    lb = torch.from_numpy(np.where(rng.random(400) < 0.7, lg.argmax(1).numpy(), rng.integers(0, C, 400)))

    class M(torch.nn.Module):  # a "model" that returns its input as logits
        def forward(self, d):
            return d, None

    loader = [(lg[i:i + 100], lb[i:i + 100]) for i in range(0, 400, 100)]
    out["temp/logits"] = lg.numpy()
    out["temp/labels"] = lb.numpy()
    out["temp/value"] = np.array(ns["fit_temperature"](M(), loader, torch.device("cpu")))
    p = os.path.join(ROOT, "tests", "golden", "head_training.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes; temperature", float(out["temp/value"]))


if __name__ == "__main__":
    main()
