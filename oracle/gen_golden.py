"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (read-only checkout at /root/reference).

Run once in the authoring container:  python -m oracle.gen_golden
The GPU box has no /root/reference; the committed fixtures travel instead.

backend/cbas.py cannot be imported as-is here (decord, h5py, matplotlib are not installed), so this script
installs minimal in-memory stand-ins for exactly those I/O modules before importing it - a numpy-backed
`decord.VideoReader`, a dict-backed `h5py.File`, an inert `matplotlib` - and then calls the reference's own
`encode_file`, `infer_file`, `ClassifierLSTMDeltas` and `Actogram` unmodified.  Only file/video I/O is faked;
every number in the fixtures is produced by reference code (and by `transformers` for the ViT, as in the
reference).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# ------------------------------------------------------------------------------------------- I/O stand-ins
_VIDEOS = {}    # path -> uint8 [n,H,W,3]
_H5 = {}        # path -> {"attrs": {}, "datasets": {name: np.ndarray}}


class _Batch:
    def __init__(self, a): self._a = a
    def asnumpy(self): return self._a


class _VideoReader:
    def __init__(self, path, ctx=None): self._a = _VIDEOS[path]
    def __len__(self): return len(self._a)
    def get_batch(self, idx): return _Batch(self._a[list(idx)])


class _Dataset:
    def __init__(self, store, name): self._s, self._n = store, name
    @property
    def shape(self): return self._s[self._n].shape
    def resize(self, size, axis=0):
        a = self._s[self._n]
        new = np.zeros((size,) + a.shape[1:], a.dtype)
        new[:min(size, len(a))] = a[:size]
        self._s[self._n] = new
    def __setitem__(self, k, v): self._s[self._n][k] = v
    def __getitem__(self, k): return self._s[self._n][k]
    def __len__(self): return len(self._s[self._n])


class _H5File:
    def __init__(self, path, mode="r"):
        if mode == "w":
            _H5[path] = {"attrs": {}, "datasets": {}}
        self._f = _H5[path]
        self.attrs = self._f["attrs"]
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def create_dataset(self, name, shape, maxshape=None, dtype="f4", chunks=None):
        self._f["datasets"][name] = np.zeros(shape, np.dtype(dtype))
        self._f.setdefault("layout", {})[name] = dict(maxshape=maxshape, dtype=str(np.dtype(dtype)), chunks=chunks)
        return _Dataset(self._f["datasets"], name)
    def __getitem__(self, name): return _Dataset(self._f["datasets"], name)
    def __contains__(self, name): return name in self._f["datasets"]
    def flush(self): pass
    def close(self): pass


def _install_stubs():
    decord = types.ModuleType("decord")
    decord.VideoReader, decord.cpu = _VideoReader, (lambda i=0: None)
    h5py = types.ModuleType("h5py")
    h5py.File = _H5File
    sys.modules["decord"], sys.modules["h5py"] = decord, h5py
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        sys.modules[name] = mock.MagicMock()
    sys.path[:0] = [REF, os.path.join(REF, "backend")]


def _os_replace_h5(src, dst):
    _H5[dst] = _H5.pop(src)


def main():
    os.makedirs(OUT, exist_ok=True)
    _install_stubs()
    import cbas  # the reference's backend/cbas.py, unmodified
    import classifier_head  # the reference's head, unmodified
    import gui_state

    from oracle import encoder as oenc
    from oracle import head as ohead

    # ---------------------------------------------------------------- 1. encode_file (real reference, ViT-B)
    frames = oenc.synthetic_frames(6, 64, 64, seed=7)
    model = oenc.build_hf_model("vitb16", seed=0, init_scale=4.0)
    with tempfile.TemporaryDirectory() as td:
        model.save_pretrained(td)
        enc = cbas.DinoEncoder(td, device="cpu")
    gui_state.proj = types.SimpleNamespace(encoder_model_identifier="synthetic:vitb16")
    _VIDEOS["/mem/clip.mp4"] = frames
    progress = []
    with mock.patch.object(cbas.os, "replace", _os_replace_h5), mock.patch.object(cbas.os.path, "exists", lambda p: False):
        out_path = cbas.encode_file(enc, "/mem/clip.mp4", progress.append)
    h5 = _H5[out_path]
    np.savez_compressed(
        os.path.join(OUT, "encode_file_vitb.npz"),
        cls=h5["datasets"]["cls"], out_path=out_path, progress=np.array(progress),
        attr_encoder=h5["attrs"]["encoder_model_identifier"], attr_schema=h5["attrs"]["schema_version"],
        layout_dtype=h5["layout"]["cls"]["dtype"], layout_chunks=np.array(h5["layout"]["cls"]["chunks"]),
        frames_seed=7, model_seed=0, init_scale=4.0)
    print("encode_file ->", out_path, h5["datasets"]["cls"].shape, h5["datasets"]["cls"].dtype)

    # ---------------------------------------------------------------- 2. heads (real reference module)
    def ref_head(sd, **kw):
        m = classifier_head.ClassifierLSTMDeltas(**kw).eval()
        missing, unexpected = m.load_state_dict(sd, strict=True), None
        return m

    # 2a. tiny hyper-parameters: everything stored
    tiny_kw = dict(in_features=24, out_features=5, seq_len=11, bottleneck_dim=16, center_window_size=2,
                   lstm_hidden_size=8)
    rng = np.random.default_rng(11)
    m = classifier_head.ClassifierLSTMDeltas(**tiny_kw).eval()
    torch.manual_seed(3)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
    x = torch.from_numpy(rng.standard_normal((7, 11, 24)).astype(np.float32))
    with torch.no_grad():
        logits, rawm = m(x)
        s, d, a = m._calculate_robust_deltas(x)
    np.savez_compressed(os.path.join(OUT, "head_tiny.npz"), x=x.numpy(), logits=logits.numpy(), rawm=rawm.numpy(),
                        smooth=s.numpy(), delta=d.numpy(), acc=a.numpy(),
                        **{"w:" + k: v.numpy() for k, v in m.state_dict().items()})

    # 2b. default hyper-parameters (768/9/31/128/64), weights from oracle.head.make_head_state(seed)
    sd = ohead.make_head_state(768, 9, 128, 64, seed=5, scale=2.0)
    m = ref_head(sd, in_features=768, out_features=9, seq_len=31)
    x = torch.from_numpy(np.random.default_rng(12).standard_normal((16, 31, 768)).astype(np.float16)).float()
    with torch.no_grad():
        logits, rawm = m(x)
    np.savez_compressed(os.path.join(OUT, "head_default.npz"), logits=logits.numpy(), rawm=rawm.numpy(),
                        state_seed=5, state_scale=2.0, x_seed=12)

    # ---------------------------------------------------------------- 3. infer_file (real reference loop)
    behaviors = ["eating", "drinking", "rearing", "climbing", "digging", "nesting", "resting", "grooming",
                 "background"]
    emb = (np.random.default_rng(13).standard_normal((130, 768)) * 1.5).astype(np.float16)
    _H5["/mem/clip_cls.h5"] = {"attrs": {}, "datasets": {"cls": emb}}
    with tempfile.TemporaryDirectory() as td:
        csvs = {}
        real_to_csv = cbas.pd.DataFrame.to_csv

        def to_csv(self, path, index=True):
            csvs[path] = (list(self.columns), self.to_numpy())
        with mock.patch.object(cbas.pd.DataFrame, "to_csv", to_csv):
            out_csv = cbas.infer_file("/mem/clip_cls.h5", m, "JonesLabModel", behaviors, 31,
                                      device=torch.device("cpu"), temperature=1.7)
    cols, probs = csvs[out_csv]
    np.savez_compressed(os.path.join(OUT, "infer_file.npz"), probs=probs.astype(np.float32), out_csv=out_csv,
                        columns=np.array(cols), emb_seed=13, emb_scale=1.5, temperature=1.7, state_seed=5,
                        state_scale=2.0)
    print("infer_file ->", out_csv, probs.shape)

    # ---------------------------------------------------------------- 4. Actogram (real reference binning)
    import pandas as pd
    rng = np.random.default_rng(14)
    lg = rng.standard_normal((5000, 9)) * 2.0
    pr = np.exp(lg) / np.exp(lg).sum(1, keepdims=True)
    pr = pr.astype(np.float32)
    df = pd.DataFrame(pr, columns=behaviors)
    acts = {}
    for b, (fps, binmin, thr) in {"eating": (10.0, 1, 0.5), "resting": (10.0, 2, 0.3), "background": (7.5, 1, 0.0)}.items():
        # the PNG rendering (cbas.py:574-644) is out of scope and matplotlib is a stand-in: skip the plot only
        with mock.patch.object(cbas, "_create_matplotlib_actogram", lambda *a, **k: None):
            a = cbas.Actogram(behavior=b, framerate=fps, start=0.0, binsize_minutes=binmin, threshold=thr,
                              lightcycle="LD", preloaded_df=df, model="JonesLabModel")
        acts[b] = (np.array(a.binned_activity, dtype=np.float64), a.binsize_frames)
    np.savez_compressed(os.path.join(OUT, "actogram.npz"), probs_seed=14, columns=np.array(behaviors),
                        **{f"bins:{b}": v[0] for b, v in acts.items()},
                        **{f"binsize:{b}": v[1] for b, v in acts.items()},
                        params=np.array([[10.0, 1, 0.5], [10.0, 2, 0.3], [7.5, 1, 0.0]]))
    print("actogram ->", {b: (len(v[0]), v[1]) for b, v in acts.items()})


if __name__ == "__main__":
    main()
