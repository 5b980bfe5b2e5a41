"""CPU tests: the oracle restatements against fixtures produced by the reference itself (oracle/gen_golden.py)
and, when the read-only checkout is present, live against the reference module."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import actogram as oact
from oracle import encoder as oenc
from oracle import head as ohead

REF_BACKEND = "/root/reference/backend"
BEHAVIORS = ["eating", "drinking", "rearing", "climbing", "digging", "nesting", "resting", "grooming", "background"]


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_head_tiny_matches_reference_fixture(golden_dir):
    g = _g(golden_dir, "head_tiny.npz")
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    x = torch.from_numpy(g["x"])
    s, d, a = ohead.robust_deltas(x, 0.3)
    np.testing.assert_allclose(s.numpy(), g["smooth"], atol=1e-6)
    np.testing.assert_allclose(d.numpy(), g["delta"], atol=1e-6)
    np.testing.assert_allclose(a.numpy(), g["acc"], atol=1e-6)
    logits, rawm = ohead.head_forward(sd, x, seq_len=11, center_window=2)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(rawm.numpy(), g["rawm"], atol=2e-5, rtol=1e-5)


def test_head_default_matches_reference_fixture(golden_dir):
    g = _g(golden_dir, "head_default.npz")
    sd = ohead.make_head_state(768, 9, 128, 64, seed=int(g["state_seed"]), scale=float(g["state_scale"]))
    x = torch.from_numpy(np.random.default_rng(int(g["x_seed"])).standard_normal((16, 31, 768)).astype(np.float16)).float()
    logits, rawm = ohead.head_forward(sd, x)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(rawm.numpy(), g["rawm"], atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("name", ["h128_l2", "h128_l1", "h64_l2_noacc", "h64_l1_noacc"])
def test_head_variants_match_reference_fixture(golden_dir, name):
    """lstm_hidden_size 128 / lstm_layers 2 (the reference's sweep) and use_acceleration=False, against outputs of
    the reference module itself (oracle/gen_golden_head_variants.py)."""
    g = _g(golden_dir, "head_variants.npz")
    hs, layers, acc, seed = (int(v) for v in g[name + ":cfg"])
    sd = ohead.make_head_state(768, 9, 128, hs, seed=seed, scale=float(g["state_scale"]), lstm_layers=layers,
                               use_acceleration=bool(acc))
    x = torch.from_numpy(np.random.default_rng(int(g["x_seed"])).standard_normal((12, 31, 768)).astype(np.float16)).float()
    logits, rawm = ohead.head_forward(sd, x)
    assert rawm.shape == (12, 2 * hs)
    np.testing.assert_allclose(logits.numpy(), g[name + ":logits"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(rawm.numpy(), g[name + ":rawm"], atol=5e-5, rtol=1e-4)


def test_delta_closed_forms():
    # closed forms of the reflect-padded differences
    x = torch.randn(3, 9, 5)
    s, d, a = ohead.robust_deltas(x, 0.3)
    np.testing.assert_allclose(d[:, 0], s[:, 0] - s[:, 1], atol=1e-6)
    np.testing.assert_allclose(d[:, 1:], s[:, 1:] - s[:, :-1], atol=1e-6)
    np.testing.assert_allclose(a[:, 0], s[:, 0] - 2 * s[:, 1] + s[:, 2], atol=1e-6)
    np.testing.assert_allclose(a[:, 1], 2 * (s[:, 1] - s[:, 0]), atol=1e-6)
    np.testing.assert_allclose(a[:, 2:], s[:, 2:] - 2 * s[:, 1:-1] + s[:, :-2], atol=1e-6)


def test_infer_windows_matches_reference_infer_file(golden_dir):
    g = _g(golden_dir, "infer_file.npz")
    sd = ohead.make_head_state(768, 9, 128, 64, seed=int(g["state_seed"]), scale=float(g["state_scale"]))
    emb = (np.random.default_rng(int(g["emb_seed"])).standard_normal((130, 768)) * float(g["emb_scale"])).astype(np.float16)
    probs = ohead.infer_windows(emb, sd, seq_len=31, temperature=float(g["temperature"]))
    assert probs.shape == g["probs"].shape == (130, 9)
    np.testing.assert_allclose(probs, g["probs"], atol=2e-6)
    assert list(g["columns"]) == BEHAVIORS
    assert str(g["out_csv"]) == "/mem/clip_JonesLabModel_outputs.csv"
    # chunking must not change anything: tiny chunks exercise the +-15 context logic
    probs_small = ohead.infer_windows(emb, sd, seq_len=31, temperature=float(g["temperature"]), chunk=37, batch=16)
    np.testing.assert_allclose(probs_small, probs, atol=2e-6)


def test_infer_windows_short_video():
    sd = ohead.make_head_state(32, 4, 16, 8, seed=1)
    emb = np.random.default_rng(0).standard_normal((5, 32)).astype(np.float16)  # shorter than half a window
    probs = ohead.infer_windows(emb, sd, seq_len=31)
    assert probs.shape == (5, 4)
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)


def test_actogram_matches_reference_fixture(golden_dir):
    g = _g(golden_dir, "actogram.npz")
    rng = np.random.default_rng(int(g["probs_seed"]))
    lg = rng.standard_normal((5000, 9)) * 2.0
    pr = (np.exp(lg) / np.exp(lg).sum(1, keepdims=True)).astype(np.float32)
    cols = list(g["columns"])
    for (fps, binmin, thr), b in zip(g["params"], ["eating", "resting", "background"]):
        bs = oact.binsize_frames(int(binmin), float(fps))
        assert bs == int(g[f"binsize:{b}"])
        bins = oact.actogram_bins(pr, cols.index(b), float(thr), bs)
        np.testing.assert_array_equal(bins, g[f"bins:{b}"].astype(np.int64))


def test_actogram_edges():
    assert oact.actogram_bins(np.zeros((0, 3), np.float32), 0, 0.5, 10).size == 0
    p = np.array([[0.6, 0.4], [0.5, 0.5], [0.2, 0.8]], np.float32)
    np.testing.assert_array_equal(oact.actogram_bins(p, 0, 0.5, 2), [1, 0])  # tie is not a max (strict <)
    np.testing.assert_array_equal(oact.actogram_bins(p, 1, 0.5, 2), [0, 1])


def test_encoder_oracle_matches_reference_encode_file(golden_dir):
    g = _g(golden_dir, "encode_file_vitb.npz")
    frames = oenc.synthetic_frames(6, 64, 64, seed=int(g["frames_seed"]))
    model = oenc.build_hf_model("vitb16", seed=int(g["model_seed"]), init_scale=float(g["init_scale"]))
    emb = oenc.encode(model, frames, mode="reference")
    ref = g["cls"].astype(np.float32)  # float16 as the reference stores it (cbas.py:420)
    assert ref.shape == (6, 768) and str(g["layout_dtype"]) == "float16"
    np.testing.assert_allclose(emb.astype(np.float16).astype(np.float32), ref, atol=2e-3, rtol=2e-3)
    assert list(g["layout_chunks"]) == [8192, 768]
    assert str(g["attr_schema"]) == "1.0"
    # frames must matter: rows are not copies of each other
    c = emb - emb.mean(0, keepdims=True)
    assert np.abs(c).max() > 0.05


def test_dinov2_oracle_matches_reference_encode_file(golden_dir):
    g = _g(golden_dir, "encode_file_dinov2reg.npz")
    side = int(g["side"])
    frames = oenc.synthetic_frames(5, side, side, seed=int(g["frames_seed"]))
    model = oenc.build_hf_dinov2_model("dinov2reg-b14", seed=int(g["model_seed"]), init_scale=float(g["init_scale"]),
                                       num_hidden_layers=int(g["layers"]))
    emb = oenc.encode(model, frames, mode="reference")
    ref = g["cls"].astype(np.float32)
    np.testing.assert_allclose(emb.astype(np.float16).astype(np.float32), ref, atol=2e-3, rtol=2e-3)
    assert np.abs(emb - emb.mean(0, keepdims=True)).max() > 0.05


def test_dinov2_state_dict_mapping_and_pos_embed():
    """Host-side packing of a Dinov2WithRegistersModel: key renaming and the position-embedding interpolation are
    the transformers implementation's own (no resampling when the grid matches)."""
    from cbas_b200 import encoder as benc
    model = oenc.build_hf_dinov2_model("dinov2reg-s14", seed=1, num_hidden_layers=2)
    cfg = benc.ViTConfig.from_hf(model.config)
    assert (cfg.family, cfg.patch_size, cfg.pos_grid, cfg.intermediate_size) == ("dinov2_with_registers", 14, 37, 1536)
    sd = benc.normalize_state_dict(dict(model.state_dict()), cfg.family)
    for k in ("embeddings.patch_embeddings.weight", "model.layer.1.attention.k_proj.bias", "model.layer.0.mlp.up_proj.weight",
              "model.layer.1.attention.o_proj.weight", "norm.weight", "embeddings.position_embeddings"):
        assert k in sd, k
    x = torch.zeros(1, 3, 256, 256)
    want = model.embeddings.interpolate_pos_encoding(torch.zeros(1, 1 + 18 * 18, 384), 256, 256)
    cls_pos, patch_pos = benc.interpolate_pos_embed(sd["embeddings.position_embeddings"], 18, 18)
    np.testing.assert_allclose(patch_pos.numpy(), want[0, 1:].numpy(), atol=1e-6)
    np.testing.assert_allclose(cls_pos.numpy(), want[0, 0].numpy(), atol=0)
    same_cls, same = benc.interpolate_pos_embed(sd["embeddings.position_embeddings"], 37, 37)
    assert torch.equal(same, sd["embeddings.position_embeddings"][0, 1:])


def test_preprocess_processor_matches_hf_processor():
    from transformers import DINOv3ViTImageProcessor
    frames = oenc.synthetic_frames(2, 96, 96, seed=3)
    proc = DINOv3ViTImageProcessor(size={"height": 64, "width": 64})
    want = proc(images=[torch.from_numpy(f).permute(2, 0, 1) for f in frames], return_tensors="pt")["pixel_values"]
    got = oenc.preprocess_processor(frames, 64)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=2e-6)


@pytest.mark.skipif(not os.path.isdir(REF_BACKEND), reason="reference checkout not present (GPU box)")
def test_head_live_against_reference_module():
    sys.path.insert(0, REF_BACKEND)
    try:
        import classifier_head
    finally:
        sys.path.remove(REF_BACKEND)
    sd = ohead.make_head_state(64, 6, 128, 64, seed=9, scale=1.5)
    m = classifier_head.ClassifierLSTMDeltas(in_features=64, out_features=6, seq_len=31).eval()
    m.load_state_dict(sd, strict=True)
    x = torch.randn(12, 31, 64)
    with torch.no_grad():
        want, want_raw = m(x)
    got, got_raw = ohead.head_forward(sd, x)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(got_raw.numpy(), want_raw.numpy(), atol=3e-5, rtol=1e-4)


@pytest.mark.parametrize("variant", [dict(), dict(lstm_hidden=128, lstm_layers=2), dict(use_acceleration=False)])
def test_eager_head_baseline_is_the_same_function(golden_dir, variant):
    """oracle/eager_gpu.EagerHead (the torch-eager / cuDNN baseline bench.py times on the GPU) against the pinned
    restatement, and its infer loop against the reference's infer_file fixture."""
    from oracle import eager_gpu
    sd = ohead.make_head_state(96, 5, 128, variant.get("lstm_hidden", 64), seed=4, scale=1.5,
                               lstm_layers=variant.get("lstm_layers", 1),
                               use_acceleration=variant.get("use_acceleration", True))
    m = eager_gpu.eager_head_from_state(sd, 96, 5)
    x = torch.randn(9, 31, 96)
    with torch.no_grad():
        got, raw = m(x)
    want, want_raw = ohead.head_forward(sd, x)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(raw.numpy(), want_raw.numpy(), atol=3e-5, rtol=1e-4)
    if not variant:
        g = _g(golden_dir, "infer_file.npz")
        sd = ohead.make_head_state(768, 9, 128, 64, seed=int(g["state_seed"]), scale=float(g["state_scale"]))
        emb = (np.random.default_rng(int(g["emb_seed"])).standard_normal((130, 768)) * float(g["emb_scale"])).astype(np.float16)
        probs = eager_gpu.eager_infer_loop(eager_gpu.eager_head_from_state(sd, 768, 9), emb, 31, "cpu",
                                           temperature=float(g["temperature"]))
        np.testing.assert_allclose(probs, g["probs"], atol=2e-6)
