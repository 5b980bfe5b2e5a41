"""Head training on the GPU: the gradient step against gradients the reference's own module and loss
produced (tests/golden/head_training.npz, oracle/gen_golden_training.py), fit_temperature against the reference's own
function, and train_lstm_model end to end on a separable synthetic task (its prediction passes and the returned
model run on the native kernels)."""
import os
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200 import training  # noqa: E402
from cbas_b200.classifier_head import ClassifierLSTMDeltas  # noqa: E402
from oracle import head as ohead  # noqa: E402


def _golden_batch(g):
    rng = np.random.default_rng(int(g["x_seed"]))
    B, C = int(g["batch"]), int(g["classes"])
    x = np.cumsum(rng.standard_normal((B, 31, 768)).astype(np.float32) * 0.3, axis=1).astype(np.float16)
    y = rng.integers(0, C, size=B)
    return torch.from_numpy(x).float().cuda(), torch.from_numpy(y).cuda()


@pytest.mark.parametrize("name,hs,layers", [("h64_l1", 64, 1), ("h128_l2", 128, 2)])
def test_gradient_step_matches_reference_gradients(golden_dir, name, hs, layers):
    """Loss terms, logits and every parameter's gradient of one training step (dropout off) vs the reference module +
    the loss of cbas.train_lstm_model, fp32 tolerance; the same weights in eval mode (native kernels) give the same
    logits as the differentiable path."""
    g = np.load(os.path.join(golden_dir, "head_training.npz"))
    x, y = _golden_batch(g)
    C = int(g["classes"])
    sd = ohead.make_head_state(768, C, 128, hs, seed=int(g["state_seed"]), scale=2.0, lstm_layers=layers)
    m = ClassifierLSTMDeltas(768, C, seq_len=31, lstm_hidden_size=hs, lstm_layers=layers)
    m.load_state_dict(sd)
    m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0, 0.5, 1.5, 1.0], device="cuda"), label_smoothing=0.1)
    logits, rawm = m(x)
    inv, pen = crit(logits, y), training.decorrelation_penalty(rawm)
    (inv + pen).backward()
    want_inv, want_pen = g[f"{name}/loss"]
    assert abs(float(inv) - want_inv) <= 2e-4 * max(1.0, abs(want_inv)) and abs(float(pen) - want_pen) <= 2e-3 * max(1.0, abs(want_pen))
    assert np.abs(logits.detach().cpu().numpy() - g[f"{name}/logits"]).max() <= 5e-4
    worst = 0.0
    for k, p in m.named_parameters():
        got = p.grad.detach().cpu().numpy().ravel()
        norm = float(g[f"{name}/grad_norm/{k}"])
        sample = got[::max(1, got.size // 257)][:257]
        # attention_head.bias has a zero gradient by construction (softmax ignores a common shift): both sides hold
        # rounding noise there, hence the absolute floors
        err = np.abs(sample - g[f"{name}/grad_sample/{k}"]).max() / max(norm / np.sqrt(got.size), 1e-5)
        worst = max(worst, abs(np.linalg.norm(got.astype(np.float64)) - norm) / max(norm, 1e-4))
        assert err <= 5e-2, f"{k}: sampled gradient entries differ by {err:.3e} of the rms entry"
    print(f"[parity] training step {name}: worst relative gradient-norm error {worst:.3e}")
    assert worst <= 2e-3
    m.eval()
    with torch.no_grad():
        native_logits, native_rawm = m(x)
    assert float((native_logits - logits.detach()).abs().max()) <= 5e-4
    assert float((native_rawm - rawm.detach()).abs().max()) <= 5e-4


def test_fit_temperature_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "head_training.npz"))
    lg, lb = torch.from_numpy(g["temp/logits"]), torch.from_numpy(g["temp/labels"])

    class Echo(torch.nn.Module):  # the fixture's "model": logits in, logits out
        def forward(self, d):
            return d, None

    loader = [(lg[i:i + 100], lb[i:i + 100]) for i in range(0, 400, 100)]
    got = training.fit_temperature(Echo(), loader, torch.device("cuda"))
    assert abs(got - float(g["temp/value"])) <= 2e-3, (got, float(g["temp/value"]))
    assert training.fit_temperature(Echo(), [], torch.device("cuda")) == 1.0


class _Windows(torch.utils.data.Dataset):
    """Separable toy task: class c = a drift along direction c plus noise; a few samples carry the failed-load label."""

    def __init__(self, n, classes, seed, broken=()):
        rng = np.random.default_rng(seed)
        dirs = np.random.default_rng(99).standard_normal((classes, 128)).astype(np.float32)
        self.y = rng.integers(0, classes, size=n)
        t = np.linspace(-1, 1, 31, dtype=np.float32)[None, :, None]
        self.x = (dirs[self.y][:, None, :] * (1.0 + 0.5 * t) + rng.standard_normal((n, 31, 128)).astype(np.float32) * 0.7)
        for i in broken:
            self.y[i] = -1

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return torch.from_numpy(self.x[i]), torch.tensor(int(self.y[i]))


def test_train_lstm_model_learns_and_returns_a_native_model():
    torch.manual_seed(3)
    behaviors = ["a", "b", "c", "d"]
    train, val = _Windows(600, 4, 1, broken=(5, 17)), _Windows(200, 4, 2)
    msgs = []
    model, reports, best = training.train_lstm_model(train, val, 31, behaviors, threading.Event(), batch_size=64, lr=3e-3,
                                                     epochs=6, device=torch.device("cuda"), patience=3,
                                                     progress_callback=msgs.append, class_weights=[1.0, 1.0, 1.0, 1.0])
    assert isinstance(model, ClassifierLSTMDeltas) and not model.training and model.in_features == 128
    assert 0 <= best < len(reports) <= 6 and any("Val F1" in m for m in msgs)
    f1 = [r.val_report["weighted avg"]["f1-score"] for r in reports]
    print(f"[parity] toy training: validation weighted F1 per epoch {[round(v, 3) for v in f1]}, best epoch {best}")
    assert max(f1) >= 0.9 and f1[best] == max(f1)
    assert reports[0].train_cm.shape == (4, 4) and int(reports[0].train_cm.sum()) == 598  # the -1 samples are dropped
    # the returned model classifies through the native kernels
    x = torch.from_numpy(val.x[:64]).cuda()
    with torch.no_grad():
        logits, _ = model.cuda()(x)
    assert (logits.argmax(1).cpu().numpy() == val.y[:64]).mean() >= 0.9
    # calibration on the validation loader
    loader = torch.utils.data.DataLoader(val, 64, collate_fn=training.collate_fn)
    t = training.fit_temperature(model, loader, torch.device("cuda"))
    assert 1e-3 < t <= 10.0


def test_train_lstm_model_edge_cases():
    behaviors = ["a", "b"]
    ev = threading.Event()
    assert training.train_lstm_model([], None, 31, behaviors, ev) == (None, None, -1)
    with pytest.raises(RuntimeError):
        training.train_lstm_model(_Windows(8, 2, 1), None, 31, behaviors, ev, device=torch.device("cpu"))
    ev.set()  # cancelled before the first epoch
    model, reports, best = training.train_lstm_model(_Windows(8, 2, 1), None, 31, behaviors, ev, device=torch.device("cuda"))
    assert model is None and reports == [] and best == -1
    # no validation set: the last epoch's weights are returned
    model, reports, best = training.train_lstm_model(_Windows(40, 2, 1), None, 31, behaviors, threading.Event(), batch_size=16,
                                                     epochs=2, device=torch.device("cuda"))
    assert model is not None and best == 1 and len(reports) == 2 and reports[0].val_report == {}
