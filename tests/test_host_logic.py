"""CPU: host-side logic of the drop-in surface - encode_file's file contract, the model bundle, the worker
threads' queue semantics and the multi-rank sharding (gloo, world_size 2).  No GPU, no compute kernels."""
import json
import os
import threading
import time
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from cbas_b200 import bundle, cbas, gui_state, parallel, store, workthreads
from cbas_b200.classifier_head import ClassifierLSTMDeltas
from cbas_b200.encoder import DinoEncoder, aa_bilinear_taps, rope_tables
from oracle import head as ohead


class _FakePipeline:
    """Stands in for pipeline.StreamedEncoder: embedding of a frame = [mean of its pixels, index, 0, ...]."""

    def __init__(self, width, fail_at=None):
        self.width, self.fail_at, self.seen = width, fail_at, 0

    def run(self, chunks, sink):
        for ch in chunks:
            if self.fail_at is not None and self.seen >= self.fail_at:
                raise RuntimeError("injected device failure")
            e = np.zeros((len(ch), self.width), np.float32)
            e[:, 0] = ch.reshape(len(ch), -1).mean(1)
            e[:, 1] = np.arange(self.seen, self.seen + len(ch))
            self.seen += len(ch)
            sink(e)
        return self.seen

    def run_reader(self, reader, video_len, sink, progress_callback=None):
        """what StreamedEncoder.run_reader does, minus the GPU: read_into a staging array chunk by chunk"""
        def chunks():
            for s in range(0, video_len, cbas.CHUNK_SIZE):
                e = min(s + cbas.CHUNK_SIZE, video_len)
                buf = np.empty((e - s,) + tuple(reader.frame_hw) + (3,), np.uint8)
                reader.read_into(s, e, buf)
                if progress_callback:
                    progress_callback(e / video_len * 100)
                yield buf
        return self.run(chunks(), sink)


def _fake_encoder(width=768):
    enc = DinoEncoder.__new__(DinoEncoder)
    nn.Module.__init__(enc)
    enc.hidden_size, enc.device, enc.preprocess = width, torch.device("cpu"), "processor"
    return enc


@pytest.fixture
def clip(tmp_path):
    frames = np.random.default_rng(0).integers(0, 256, (1100, 32, 32, 3), dtype=np.uint8)
    p = str(tmp_path / "cam1_00001.npy")
    np.save(p, frames)
    return p, frames


def test_encode_file_contract(monkeypatch, clip):
    path, frames = clip
    pipe = _FakePipeline(768)
    monkeypatch.setattr(cbas, "_make_pipeline", lambda enc, hw, planes=False: pipe)
    monkeypatch.setattr(gui_state, "proj", types.SimpleNamespace(encoder_model_identifier="facebook/dinov3-vitb16-pretrain-lvd1689m"))
    progress = []
    out = cbas.encode_file(_fake_encoder(), path, progress.append)
    assert out == path[:-4] + "_cls.h5" and os.path.exists(out) and not os.path.exists(out + ".tmp")
    np.testing.assert_allclose(progress, [512 / 1100 * 100, 1024 / 1100 * 100, 100.0])  # once per 512-frame chunk
    with store.EmbeddingReader(out) as r:
        assert r.shape == (1100, 768)
        assert r.attrs == {"encoder_model_identifier": "facebook/dinov3-vitb16-pretrain-lvd1689m", "schema_version": "1.0"}
        e = r.read(0, 1100)
    assert e.dtype == np.float16
    np.testing.assert_array_equal(e[:, 1], np.arange(1100).astype(np.float16))
    np.testing.assert_allclose(e[:, 0], frames.reshape(1100, -1).mean(1), rtol=2e-3)


def test_encode_file_unstamped_without_project(monkeypatch, clip):
    path, _ = clip
    monkeypatch.setattr(cbas, "_make_pipeline", lambda enc, hw, planes=False: _FakePipeline(384))
    monkeypatch.setattr(gui_state, "proj", None)
    out = cbas.encode_file(_fake_encoder(384), path)
    with store.EmbeddingReader(out) as r:
        assert r.attrs == {} and r.shape == (1100, 384)  # cbas.py:414: stamps only when a project is loaded


def test_encode_file_failure_removes_tmp_and_raises(monkeypatch, clip):
    path, _ = clip
    monkeypatch.setattr(cbas, "_make_pipeline", lambda enc, hw, planes=False: _FakePipeline(768, fail_at=512))
    with pytest.raises(RuntimeError, match="injected"):
        cbas.encode_file(_fake_encoder(), path)
    d = os.path.dirname(path)
    assert [f for f in os.listdir(d) if f.endswith((".h5", ".tmp"))] == []


def test_encode_file_empty_and_bad_inputs(monkeypatch, tmp_path):
    monkeypatch.setattr(cbas, "_make_pipeline", lambda enc, hw, planes=False: _FakePipeline(768))
    p = str(tmp_path / "empty.npy")
    np.save(p, np.zeros((0, 32, 32, 3), np.uint8))
    assert cbas.encode_file(_fake_encoder(), p) is None            # zero frames -> None (cbas.py:405-407)
    with pytest.raises(Exception):
        cbas.encode_file(_fake_encoder(), str(tmp_path / "missing.mp4"))  # decode errors propagate
    with pytest.raises(TypeError):
        cbas.encode_file(nn.Linear(2, 2), p)


def test_video_reader_decodes_mp4(tmp_path):
    cv2 = pytest.importorskip("cv2")
    p = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (64, 48))
    if not vw.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    for i in range(12):
        vw.write(np.full((48, 64, 3), (i * 20, 10, 255 - i * 20), np.uint8))  # BGR
    vw.release()
    r = cbas.VideoReader(p)
    assert len(r) == 12
    b = r.get_batch(range(3, 7))
    assert b.shape == (4, 48, 64, 3) and b.dtype == np.uint8
    assert abs(int(b[0, 10, 10, 2]) - 60) < 12 and abs(int(b[0, 10, 10, 0]) - 195) < 12  # RGB order
    r.close()


def test_parallel_video_reader_matches_sequential_decode(tmp_path):
    """Chunks decoded ahead by worker processes into shared memory are the frames the in-thread reader returns,
    in order, including the short last chunk; out-of-order requests and broken files are reported."""
    cv2 = pytest.importorskip("cv2")
    from cbas_b200.decode import ParallelVideoReader
    p = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (64, 48))
    if not vw.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    rng = np.random.default_rng(0)
    base = rng.integers(0, 255, (48, 64, 3), dtype=np.uint8)
    for i in range(141):
        vw.write(np.roll(base, 3 * i, axis=1))
    vw.release()
    seq = cbas.VideoReader(p)
    want = seq.get_batch(range(0, len(seq)))
    seq.close()
    par = ParallelVideoReader(p, workers=3, chunk=32)
    try:
        assert len(par) == 141 and par.n_chunks == 5
        got = []
        for i in range(0, 141, 32):
            b = par.get_batch(range(i, min(i + 32, 141)))
            got.append(np.array(b))  # copy: the view is recycled two requests later
        assert np.array_equal(np.concatenate(got), want)
        with pytest.raises(ValueError):
            par.get_batch(range(0, 32))  # chunks are served once, in order
    finally:
        par.close()
    with pytest.raises(FileNotFoundError):
        ParallelVideoReader(str(tmp_path / "missing.mp4"))


def test_cloned_readers_decode_whole_chunks_exactly(tmp_path):
    """pipeline.run_reader gives each of several in-process decoders (VideoReader.clone) every R-th chunk: a clone that
    seeks to its chunk and decodes it must produce the frames the single sequential reader produces, RGB and
    green-plane form, and a frame-range view (one video on several GPUs) clones into the same range."""
    cv2 = pytest.importorskip("cv2")
    p = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (64, 48))
    if not vw.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    base = np.random.default_rng(1).integers(0, 255, (48, 64, 3), dtype=np.uint8)
    for i in range(150):
        vw.write(np.roll(base, 2 * i, axis=0))
    vw.release()
    seq = cbas.VideoReader(p)
    want = seq.get_batch(range(0, len(seq)))
    assert len(seq) == 150 and seq.parallel_readers >= 1
    readers = [seq, seq.clone(), seq.clone()]
    starts = list(range(0, 150, 32))
    got = np.zeros_like(want)
    green = np.zeros(want.shape[:3], np.uint8)
    for j, r in enumerate(readers):          # reader j owns chunks j, j+3, ... like run_reader's threads
        for k in range(j, len(starts), 3):
            s, e = starts[k], min(starts[k] + 32, 150)
            r.read_into(s, e, got[s:e])
    for j, r in enumerate(reversed(readers)):  # and again with other owners, green plane only (REFERENCE mode)
        for k in range(j, len(starts), 3):
            s, e = starts[k], min(starts[k] + 32, 150)
            r.read_into(s, e, green[s:e])
    assert np.array_equal(got, want) and np.array_equal(green, want[..., 1])
    span = cbas._SpanReader(seq, 40, 110)
    twin = span.clone()
    buf = np.zeros((30, 48, 64, 3), np.uint8)
    twin.read_into(10, 40, buf)
    assert len(twin) == 70 and np.array_equal(buf, want[50:80])
    for r in readers[1:]:
        r.close()
    twin.close()
    seq.close()


def test_infer_file_swallows_errors_and_returns_none(tmp_path, capsys):
    out = cbas.infer_file(str(tmp_path / "nope_cls.h5"), nn.Linear(1, 1), "m", ["a"], 31, device="cpu")
    assert out is None and "Error during buffered inference" in capsys.readouterr().out


def test_bundle_round_trip_and_fallbacks(tmp_path):
    behaviors = ["eating", "drinking", "resting"]
    sd = ohead.make_head_state(768, 3, 128, 64, seed=2)
    m = ClassifierLSTMDeltas(768, 3, seq_len=31)
    m.load_state_dict(sd)
    d = str(tmp_path / "JonesLabModel")
    bundle.save_model_bundle(d, m, "JonesLabModel", behaviors, 31, "facebook/dinov3-vitb16-pretrain-lvd1689m", 1.37)
    assert sorted(os.listdir(d)) == ["config.yaml", "model.pth", "model_meta.json"]
    meta = json.load(open(os.path.join(d, "model_meta.json")))
    assert meta["model_bundle_schema"] == "1.0" and meta["head_architecture_version"] == "ClassifierLSTMDeltas"
    assert meta["hyperparameters"] == {"behaviors": behaviors, "seq_len": 31, "use_acceleration": True,
                                       "lstm_hidden_size": 64, "lstm_layers": 1}
    assert meta["calibration"]["temperature"] == pytest.approx(1.37)
    m2, meta2 = bundle.load_model_bundle(d, "facebook/dinov3-vitb16-pretrain-lvd1689m", device="cpu")
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])
    assert not m2.training and meta2["hyperparameters"]["behaviors"] == behaviors
    # hyper-parameters missing from the metadata are inferred from the weight shapes (workthreads.py:416-425)
    del meta["hyperparameters"]["lstm_hidden_size"], meta["hyperparameters"]["lstm_layers"], meta["hyperparameters"]["behaviors"]
    json.dump(meta, open(os.path.join(d, "model_meta.json"), "w"))
    m3, meta3 = bundle.load_model_bundle(d, None, device="cpu")
    assert meta3["hyperparameters"]["lstm_hidden_size"] == 64 and meta3["hyperparameters"]["lstm_layers"] == 1
    assert meta3["hyperparameters"]["behaviors"] == behaviors  # from config.yaml
    with pytest.raises(bundle.EncoderMismatch):
        bundle.load_model_bundle(d, "facebook/dinov2-with-registers-base", device="cpu")
    os.remove(os.path.join(d, "model_meta.json"))  # legacy bundle: no metadata -> legacy architecture
    with pytest.raises(NotImplementedError):
        bundle.load_model_bundle(d, None, device="cpu")


def test_head_is_cuda_only_in_both_modes():
    m = ClassifierLSTMDeltas(768, 9)
    assert not m.training and not any(p.requires_grad for p in m.parameters())
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 31, 768))          # no CPU fallback
    m.train()                                # the differentiable path (cbas_b200.training) is GPU-only as well
    assert all(p.requires_grad for p in m.parameters())
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 31, 768))
    m.eval()
    assert not any(p.requires_grad for p in m.parameters())


def test_time_operators_reproduce_the_three_streams():
    """The [T,T] operators of the differentiable head path against a direct statement of classifier_head.py:100-116
    (EMA, reflected two-frame head, first and second differences), for an ordinary and a two-frame window."""
    for T in (31, 7, 2, 1):
        m = ClassifierLSTMDeltas(8, 3, seq_len=T)
        S, D1, D2 = m._time_operators(T, torch.device("cpu"), torch.float64)
        x = torch.randn(T, 8, dtype=torch.float64)
        sm = x.clone()
        for t in range(1, T):
            sm[t] = sm[t - 1] + m.ema_alpha * (x[t] - sm[t - 1])
        head = [sm[2], sm[1]] if T >= 3 else [sm[0], sm[0]]
        padded = torch.stack(head + list(sm))
        dx = padded[1:] - padded[:-1]
        assert torch.allclose(S @ x, sm, atol=1e-12)
        assert torch.allclose(D1 @ x, dx[1:], atol=1e-12)
        assert torch.allclose(D2 @ x, dx[1:] - dx[:-1], atol=1e-12)


def test_encode_thread_queue_semantics(monkeypatch):
    done, fail = [], {"b.mp4"}

    def fake_encode(encoder, path, cb=None):
        if cb:
            cb(100.0)
        if os.path.basename(path) in fail:
            raise RuntimeError("bad video")
        done.append(path)
        return path.replace(".mp4", "_cls.h5")

    monkeypatch.setattr(cbas, "encode_file", fake_encode)
    monkeypatch.setattr(gui_state, "dino_encoder", object())
    monkeypatch.setattr(gui_state, "live_inference_model_name", "JonesLabModel")
    gui_state.encode_tasks[:] = ["/v/a.mp4", "/v/b.mp4", "/v/c.mp4"]
    gui_state.classify_tasks[:] = []
    t = workthreads.EncodeThread("cpu", poll_seconds=0.01)
    t.start()
    for _ in range(500):
        if not gui_state.encode_tasks and len(done) == 2:
            break
        time.sleep(0.01)
    t.stop()
    t.join(2)
    assert done == ["/v/a.mp4", "/v/c.mp4"]                                  # FIFO, the failing file is skipped
    assert gui_state.classify_tasks == ["/v/a_cls.h5", "/v/c_cls.h5"]      # chained to live inference
    gui_state.classify_tasks[:] = []


def test_partitioning_is_balanced_and_deterministic():
    paths = [f"cam{c}/seg{s:03d}.mp4" for c in range(8) for s in range(9)]
    costs = [6000 + (i * 37) % 500 for i in range(len(paths))]
    shards = parallel.partition_videos(paths, costs, 8)
    assert sorted(p for s in shards for p in s) == sorted(paths)
    loads = [sum(costs[paths.index(p)] for p in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)
    assert shards == parallel.partition_videos(list(reversed(paths)), list(reversed(costs)), 8)
    spans = parallel.split_frame_range(18000, 8, halo=15)
    assert spans[0].start == 0 and spans[-1].stop == 18000 and spans[1].start == 2250 - 15


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    paths = [f"cam{c}/seg{s}.mp4" for c in range(3) for s in range(5)]
    costs = [100 + 7 * i for i in range(len(paths))]
    mine = parallel.partition_videos(paths, costs, world)[rank]
    n_bins = {"cam0": 4, "cam1": 4, "cam2": 6}
    local = {}
    for p in mine:  # a fake per-video actogram: bin k gets (index of the video) + k
        cam, i = p.split("/")[0], paths.index(p)
        local.setdefault(cam, torch.zeros(n_bins[cam], dtype=torch.int64))
        local[cam] += torch.arange(n_bins[cam]) + i
    total = parallel.allreduce_bins(local, n_bins)
    q.put((rank, mine, {k: v.tolist() for k, v in total.items()}))
    dist.destroy_process_group()


def test_two_rank_sharding_and_bin_reduction_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    paths = [f"cam{c}/seg{s}.mp4" for c in range(3) for s in range(5)]
    got = {r: (mine, tot) for r, mine, tot in res}
    assert sorted(got[0][0] + got[1][0]) == sorted(paths) and not set(got[0][0]) & set(got[1][0])
    want = {}
    for i, p in enumerate(paths):
        cam = p.split("/")[0]
        n = {"cam0": 4, "cam1": 4, "cam2": 6}[cam]
        want[cam] = [a + b for a, b in zip(want.get(cam, [0] * n), [k + i for k in range(n)])]
    assert got[0][1] == got[1][1] == want


def test_host_side_tables():
    xmin, w = aa_bilinear_taps(256, 224)
    assert xmin.shape == (224,) and w.shape[0] == 224 and np.allclose(w.sum(1), 1.0, atol=1e-6)
    x = torch.rand(1, 1, 256, 256)
    want = torch.nn.functional.interpolate(x, size=(224, 224), mode="bilinear", antialias=True)[0, 0]
    ymin, wy = aa_bilinear_taps(256, 224)
    rows = torch.stack([sum(float(wy[i, j]) * x[0, 0, min(int(ymin[i]) + j, 255)] for j in range(wy.shape[1])) for i in range(224)])
    got = torch.stack([sum(float(w[i, j]) * rows[:, min(int(xmin[i]) + j, 255)] for j in range(w.shape[1])) for i in range(224)], dim=1)
    assert (got - want).abs().max() < 2e-6
    c, s = rope_tables(14, 14)
    assert c.shape == (196, 32) and torch.allclose(c * c + s * s, torch.ones_like(c), atol=1e-6)
