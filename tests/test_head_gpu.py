"""LSTM head + actogram on the GPU vs the CPU oracle and the fixtures produced by the reference itself.
Gates (BASELINE.json): probabilities within 1e-3, argmax agreement >= 99.9 %; actogram bins exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200.classifier_head import ClassifierLSTMDeltas, actogram_bins  # noqa: E402
from oracle import actogram as oact  # noqa: E402
from oracle import head as ohead  # noqa: E402

BEHAVIORS = ["eating", "drinking", "rearing", "climbing", "digging", "nesting", "resting", "grooming", "background"]


def _head(sd, **kw):
    m = ClassifierLSTMDeltas(**kw)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.to("cuda").eval()


def _check_probs(got, want, tag):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    d = np.abs(got - want).max()
    agree = (got.argmax(1) == want.argmax(1)).mean()
    print(f"[parity] {tag}: max|dp| {d:.3e}  argmax agreement {agree * 100:.3f}%  ({len(got)} frames)")
    assert d <= 1e-3, f"{tag}: probabilities differ by {d}"
    assert agree >= 0.999, f"{tag}: argmax agreement {agree}"


def test_infer_file_fixture_from_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "infer_file.npz"))
    sd = ohead.make_head_state(768, 9, 128, 64, seed=int(g["state_seed"]), scale=float(g["state_scale"]))
    emb = (np.random.default_rng(int(g["emb_seed"])).standard_normal((130, 768)) * float(g["emb_scale"])).astype(np.float16)
    head = _head(sd, in_features=768, out_features=9, seq_len=31)
    probs = head.infer_embeddings(torch.from_numpy(emb).cuda(), temperature=float(g["temperature"])).cpu().numpy()
    _check_probs(probs, g["probs"], "infer_file fixture (reference run)")


def test_forward_windows_fixture_from_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "head_default.npz"))
    sd = ohead.make_head_state(768, 9, 128, 64, seed=int(g["state_seed"]), scale=float(g["state_scale"]))
    x = torch.from_numpy(np.random.default_rng(int(g["x_seed"])).standard_normal((16, 31, 768)).astype(np.float16)).float()
    head = _head(sd, in_features=768, out_features=9, seq_len=31)
    logits, rawm = head(x.cuda())
    assert logits.shape == (16, 9) and rawm.shape == (16, 128)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=1e-3, rtol=1e-3)
    np.testing.assert_allclose(rawm.cpu().numpy(), g["rawm"], atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("n,F,C,seq,scale,temp", [
    (700, 768, 9, 31, 2.0, 1.0), (300, 384, 5, 31, 1.0, 0.5), (257, 1024, 12, 63, 1.5, 2.5), (40, 768, 9, 31, 3.0, 1e-4),
    (9000, 768, 9, 31, 2.0, 1.3),
])
def test_infer_vs_oracle(n, F, C, seq, scale, temp):
    sd = ohead.make_head_state(F, C, 128, 64, seed=n, scale=scale)
    emb = (np.random.default_rng(n + 1).standard_normal((n, F)) * 1.2).astype(np.float16)
    # a slowly varying component so neighbouring frames correlate like real embeddings
    emb = (emb.astype(np.float32) + 2.0 * np.sin(np.arange(n)[:, None] / 17.0 + np.arange(F)[None, :])).astype(np.float16)
    head = _head(sd, in_features=F, out_features=C, seq_len=seq)
    probs, logits = head.infer_embeddings(torch.from_numpy(emb).cuda(), temperature=temp, return_logits=True)
    # the oracle takes seconds per thousand windows: check every frame for small n, a spread sample for large n
    if n <= 1000:
        idx = np.arange(n)
        want, want_logits = ohead.infer_windows(emb, sd, seq_len=seq, temperature=temp, return_logits=True)
    else:
        idx = np.unique(np.concatenate([np.arange(0, 40), np.arange(4080, 4120), np.arange(8170, 8210), np.arange(n - 40, n)]))
        half = seq // 2
        pad = np.concatenate([np.repeat(emb[:1], half, 0), emb, np.repeat(emb[-1:], half, 0)])
        win = torch.from_numpy(np.stack([pad[i:i + seq] for i in idx])).float()
        with torch.no_grad():
            want_logits, _ = ohead.head_forward(sd, win, seq_len=seq)
        want = torch.softmax(want_logits / max(1e-3, temp), dim=1).numpy()
        want_logits = want_logits.numpy()
    np.testing.assert_allclose(logits.cpu().numpy()[idx], want_logits, atol=2e-4, rtol=2e-4)
    _check_probs(probs.cpu().numpy()[idx], want, f"head F{F} C{C} T{seq} n{n} temp{temp}")


@pytest.mark.parametrize("name", ["h128_l2", "h128_l1", "h64_l2_noacc", "h64_l1_noacc"])
def test_head_variants_fixture_from_reference(golden_dir, name):
    """Non-default heads (lstm_hidden_size 128, two LSTM layers, no acceleration stream): forward(x) against the
    reference module's own outputs, and the stride-1 inference path against the oracle."""
    g = np.load(os.path.join(golden_dir, "head_variants.npz"))
    hs, layers, acc, seed = (int(v) for v in g[name + ":cfg"])
    sd = ohead.make_head_state(768, 9, 128, hs, seed=seed, scale=float(g["state_scale"]), lstm_layers=layers,
                               use_acceleration=bool(acc))
    x = torch.from_numpy(np.random.default_rng(int(g["x_seed"])).standard_normal((12, 31, 768)).astype(np.float16)).float()
    head = _head(sd, in_features=768, out_features=9, seq_len=31, lstm_hidden_size=hs, lstm_layers=layers,
                 use_acceleration=bool(acc))
    logits, rawm = head(x.cuda())
    assert logits.shape == (12, 9) and rawm.shape == (12, 2 * hs)
    np.testing.assert_allclose(logits.cpu().numpy(), g[name + ":logits"], atol=1e-3, rtol=1e-3)
    np.testing.assert_allclose(rawm.cpu().numpy(), g[name + ":rawm"], atol=1e-3, rtol=1e-3)
    n = 333
    emb = (np.random.default_rng(seed).standard_normal((n, 768)) * 1.2).astype(np.float16)
    probs = head.infer_embeddings(torch.from_numpy(emb).cuda(), temperature=1.7).cpu().numpy()
    want = ohead.infer_windows(emb, sd, seq_len=31, temperature=1.7)
    _check_probs(probs, want, f"head variant {name}")


@pytest.mark.parametrize("seq,cw,C,F,hs,layers", [(95, 5, 9, 768, 64, 1), (63, 9, 32, 384, 128, 1), (3, 5, 1, 64, 64, 2),
                                                  (5, 1, 2, 128, 128, 2), (31, 0, 9, 768, 64, 1), (11, 20, 3, 256, 64, 1)])
def test_head_hyperparameter_corners(seq, cw, C, F, hs, layers):
    """Corners of the head's configuration space: the sweep's long windows (63 / 95, sweep_runner.py:110), the shortest
    legal window, a centre window wider than the sequence (clipped to it) and of a single frame, 1 and 32 behaviours."""
    sd = ohead.make_head_state(F, C, 128, hs, seed=seq + C, scale=1.5, lstm_layers=layers)
    n = 150
    emb = (np.random.default_rng(seq).standard_normal((n, F)) * 1.1).astype(np.float16)
    head = _head(sd, in_features=F, out_features=C, seq_len=seq, center_window_size=cw, lstm_hidden_size=hs,
                 lstm_layers=layers)
    probs, logits = head.infer_embeddings(torch.from_numpy(emb).cuda(), temperature=0.8, return_logits=True)
    want, want_logits = ohead.infer_windows(emb, sd, seq_len=seq, temperature=0.8, return_logits=True, center_window=cw)
    np.testing.assert_allclose(logits.cpu().numpy(), want_logits, atol=3e-4, rtol=3e-4)
    _check_probs(probs.cpu().numpy(), want, f"head corner T{seq} cw{cw} C{C} F{F} Hs{hs} L{layers}")


@pytest.mark.parametrize("n", [1, 5, 15, 16, 31])
def test_short_videos_are_all_padding(n):
    sd = ohead.make_head_state(768, 9, 128, 64, seed=3)
    emb = np.random.default_rng(n).standard_normal((n, 768)).astype(np.float16)
    head = _head(sd, in_features=768, out_features=9, seq_len=31)
    probs = head.infer_embeddings(torch.from_numpy(emb).cuda()).cpu().numpy()
    want = ohead.infer_windows(emb, sd, seq_len=31)
    _check_probs(probs, want, f"short video n={n}")


def test_empty_video():
    sd = ohead.make_head_state(768, 9, 128, 64, seed=3)
    head = _head(sd, in_features=768, out_features=9, seq_len=31)
    probs = head.infer_embeddings(torch.zeros(0, 768, dtype=torch.float16, device="cuda"))
    assert probs.shape == (0, 9)


def test_context_invariance_at_full_size():
    """BASELINE config 4 size (1M frames): rows are probabilities, and a frame's output depends only on its
    +-15 neighbours - the slice [a-15, b+15) alone must reproduce frames [a, b) of the full run."""
    n = 1_000_000
    sd = ohead.make_head_state(768, 9, 128, 64, seed=7, scale=2.0)
    head = _head(sd, in_features=768, out_features=9, seq_len=31)
    g = torch.Generator(device="cuda").manual_seed(0)
    emb = torch.randn(n, 768, device="cuda", generator=g).half()
    probs = head.infer_embeddings(emb)
    assert torch.isfinite(probs).all()
    assert (probs.sum(1) - 1).abs().max() < 1e-5
    # one range inside a chunk of the window stages, one across a chunk boundary (chunks of 2 * 148 * 128 = 37 888 windows:
    # different chunks, different 128-window tiles of the recurrence kernel, same numbers)
    for a, b in ((500_000, 500_400), (3 * 37_888 - 200, 3 * 37_888 + 200)):
        sub = head.infer_embeddings(emb[a - 15:b + 15].contiguous())
        assert (sub[15:-15] - probs[a:b]).abs().max() < 1e-6


def test_actogram_fixture_from_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "actogram.npz"))
    rng = np.random.default_rng(int(g["probs_seed"]))
    lg = rng.standard_normal((5000, 9)) * 2.0
    pr = (np.exp(lg) / np.exp(lg).sum(1, keepdims=True)).astype(np.float32)
    cols = list(g["columns"])
    prd = torch.from_numpy(pr).cuda()
    for (fps, binmin, thr), b in zip(g["params"], ["eating", "resting", "background"]):
        bs = oact.binsize_frames(int(binmin), float(fps))
        bins = actogram_bins(prd, cols.index(b), float(thr), bs).cpu().numpy()
        np.testing.assert_array_equal(bins, g[f"bins:{b}"].astype(np.int64))


def test_actogram_vs_oracle_large_and_edges():
    rng = np.random.default_rng(5)
    pr = rng.dirichlet(np.ones(9) * 0.3, size=200_003).astype(np.float32)
    pr[::7] = pr[::7][:, ::-1]
    pr[100:200, 2] = pr[100:200, 4]  # exact ties are not a maximum (strict <)
    prd = torch.from_numpy(pr).cuda()
    for b, thr, bs in [(0, 0.5, 6000), (4, 0.2, 1), (8, 0.0, 200_003), (2, 0.9, 250_000)]:
        got = actogram_bins(prd, b, thr, bs).cpu().numpy()
        np.testing.assert_array_equal(got, oact.actogram_bins(pr, b, thr, bs))
    assert actogram_bins(torch.zeros(0, 9, device="cuda"), 0, 0.5, 10).numel() == 0
    # single behaviour: the reference's max over zero other columns is NaN, `NaN < p` is False, so the product is 0
    one = torch.rand(50, 1, device="cuda") * 0.5 + 0.5
    for thr in (0.5, 0.0):  # 0 >= 0.5 never; 0 >= 0.0 always (what pandas/numpy give in the reference)
        np.testing.assert_array_equal(actogram_bins(one, 0, thr, 10).cpu().numpy(),
                                      oact.actogram_bins(one.cpu().numpy(), 0, thr, 10))


def test_actogram_float64_tables_compare_like_the_reference():
    """Actogram reads the `_outputs.csv` tables as float64 and compares in float64 (cbas.py:989-993): a probability
    just under the threshold that float32 would round up onto it must not count, ties of the other-behaviour maximum
    are decided on the float64 values, NaN cells are skipped by the row maximum."""
    thr = 0.7
    below = np.nextafter(np.float64(np.float32(thr)), 0.0)        # float32(thr) rounds ABOVE 0.7; one float64 step below it
    rows = np.array([[0.1, np.float64(np.float32(thr)), 0.2],     # >= 0.7 in float64: counts
                     [0.1, 0.7 - 1e-12, 0.2],                     # float32 rounds it to float32(0.7) >= thr32: must NOT count
                     [0.1, below, 0.2],
                     [0.8, 0.8 + 1e-13, 0.0],                     # is_max decided in float64 (float32 sees a tie): counts
                     [np.nan, 0.9, 0.3],                          # NaN skipped by max(axis=1): counts
                     [0.95, 0.9, 0.0]], dtype=np.float64)
    want = oact.actogram_bins(rows, 1, thr, 2)
    got = actogram_bins(torch.from_numpy(rows).cuda(), 1, thr, 2).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert want.tolist() == [0, 1, 1]
    # the same table through float32 WOULD count row 1 (the rounding the float64 path exists to avoid)
    got32 = actogram_bins(torch.from_numpy(rows.astype(np.float32)).cuda(), 1, thr, 2).cpu().numpy()
    assert got32.tolist() != want.tolist()
