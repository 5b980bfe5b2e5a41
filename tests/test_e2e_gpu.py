"""The whole drop-in path on the GPU through the public API: encode_file -> `_cls.h5` -> infer_file -> CSV ->
Actogram, against the CPU oracle on the same synthetic clip, plus the worker-thread chain."""
import os
import threading
import time
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200 import bundle, cbas, gui_state, store, workthreads  # noqa: E402
from cbas_b200.classifier_head import ClassifierLSTMDeltas  # noqa: E402
from cbas_b200.encoder import DinoEncoder  # noqa: E402
from oracle import actogram as oact  # noqa: E402
from oracle import encoder as oenc  # noqa: E402
from oracle import head as ohead  # noqa: E402

BEHAVIORS = ["eating", "drinking", "rearing", "climbing", "digging", "nesting", "resting", "grooming", "background"]


def test_encode_infer_actogram_chain(tmp_path, monkeypatch):
    model = oenc.build_hf_model("vitb16", seed=0, init_scale=3.0)
    frames = oenc.synthetic_frames(70, 64, 64, seed=11)
    clip = str(tmp_path / "cam1_00001.npy")
    np.save(clip, frames)
    monkeypatch.setattr(gui_state, "proj", types.SimpleNamespace(encoder_model_identifier="synthetic:vitb16"))
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=32)
    monkeypatch.setattr(cbas, "CHUNK_SIZE", 32)  # three chunks (32+32+6) through the double-buffered pipeline
    progress = []
    out = cbas.encode_file(enc, clip, progress.append)
    assert out == str(tmp_path / "cam1_00001_cls.h5") and progress[-1] == 100.0 and len(progress) == 3
    with store.EmbeddingReader(out) as r:
        assert r.shape == (70, 768) and r.attrs["encoder_model_identifier"] == "synthetic:vitb16"
        emb = r.read(0, 70)
    want = oenc.encode(model, frames, mode="reference")
    rel = np.abs(emb.astype(np.float32) - want).max() / np.abs(want).max()
    cos = (emb.astype(np.float32) * want).sum(1) / (np.linalg.norm(emb.astype(np.float32), axis=1) * np.linalg.norm(want, axis=1))
    print(f"[parity] encode_file -> h5: min cosine {cos.min():.6f} max|d|/max|ref| {rel:.3e}")
    assert cos.min() >= 0.999 and rel <= 2e-2

    sd = ohead.make_head_state(768, 9, 128, 64, seed=4, scale=2.0)
    head = ClassifierLSTMDeltas(768, 9, seq_len=31)
    head.load_state_dict(sd)
    csv = cbas.infer_file(out, head, "JonesLabModel", BEHAVIORS, 31, device=torch.device("cuda"), temperature=1.4)
    assert csv == str(tmp_path / "cam1_00001_JonesLabModel_outputs.csv")
    import pandas as pd
    df = pd.read_csv(csv)
    assert list(df.columns) == BEHAVIORS and len(df) == 70
    want_p = ohead.infer_windows(emb, sd, seq_len=31, temperature=1.4)  # the oracle head on the SAME stored f16 rows
    assert np.abs(df.to_numpy() - want_p).max() <= 1e-3
    assert (df.to_numpy().argmax(1) == want_p.argmax(1)).mean() >= 0.999

    act = cbas.Actogram("resting", framerate=0.1, start=0.0, binsize_minutes=2, threshold=0.1, lightcycle="LD",
                        preloaded_df=df, model="JonesLabModel")
    bs = oact.binsize_frames(2, 0.1)
    assert act.binsize_frames == bs == 12
    np.testing.assert_array_equal(np.array(act.binned_activity, np.int64),
                                  oact.actogram_bins(df.to_numpy(dtype=np.float32), BEHAVIORS.index("resting"), 0.1, bs))
    act2 = cbas.Actogram("resting", 0.1, 0.0, 2, 0.1, "LD", directory=str(tmp_path), model="JonesLabModel")
    assert act2.binned_activity == act.binned_activity


def test_encode_file_from_mp4_with_parallel_decode(tmp_path, monkeypatch):
    """A real container file through encode_file: the in-thread decoder (the reference's arrangement) and the
    process-pool decoder must produce the same `_cls.h5`, and both must match the oracle run on the decoded frames."""
    cv2 = pytest.importorskip("cv2")
    clip = str(tmp_path / "cam2_00001.mp4")
    vw = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (64, 64))
    if not vw.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    for f in oenc.synthetic_frames(75, 64, 64, seed=21):
        vw.write(f[:, :, ::-1].copy())  # RGB -> BGR
    vw.release()
    monkeypatch.setattr(gui_state, "proj", None)
    monkeypatch.setattr(cbas, "CHUNK_SIZE", 32)
    enc = DinoEncoder("synthetic:vits16@3", "cuda", max_frames=32)
    outs = {}
    for workers in (0, 3):
        monkeypatch.setattr(cbas, "DECODE_WORKERS", workers)
        out = cbas.encode_file(enc, clip)
        with store.EmbeddingReader(out) as r:
            assert r.shape == (75, 384)
            outs[workers] = r.read(0, 75)
        os.remove(out)
    assert np.array_equal(outs[0], outs[3])  # same frames, same kernels -> identical f16 rows
    reader = cbas.VideoReader(clip)
    decoded = reader.get_batch(range(75))
    reader.close()
    direct = enc.encode_u8(torch.from_numpy(decoded).cuda()).cpu().numpy().astype(np.float16)
    assert np.abs(outs[0].astype(np.float32) - direct.astype(np.float32)).max() <= 2e-3 * np.abs(direct.astype(np.float32)).max()


def test_baseline_config0_vits16_256px_then_head(tmp_path, monkeypatch):
    """BASELINE configs[0] (the reference's CPU-runnable case): DINOv3 ViT-S/16 on a 256x256 clip in the reference's
    preprocessing (green/255, native resolution: 261 tokens per frame -> the key-split attention kernel), batches of
    32, then the LSTM head with in_features = 384 on the stored f16 rows.  The clip is 48 frames instead of 300 so that
    the fp32 CPU oracle finishes in seconds; nothing in the path depends on the clip length."""
    model = oenc.build_hf_model("vits16", seed=3, init_scale=3.0)
    frames = oenc.synthetic_frames(48, 256, 256, seed=31)
    clip = str(tmp_path / "cam3_00001.npy")
    np.save(clip, frames)
    monkeypatch.setattr(gui_state, "proj", None)
    monkeypatch.setattr(cbas, "CHUNK_SIZE", 32)
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=32)
    out = cbas.encode_file(enc, clip)
    with store.EmbeddingReader(out) as r:
        assert r.shape == (48, 384)
        emb = r.read(0, 48)
    want = oenc.encode(model, frames, mode="reference", batch=32)
    e32 = emb.astype(np.float32)
    rel = np.abs(e32 - want).max() / np.abs(want).max()
    cos = (e32 * want).sum(1) / (np.linalg.norm(e32, axis=1) * np.linalg.norm(want, axis=1))
    print(f"[parity] configs[0] ViT-S/16 @256: min cosine {cos.min():.6f} max|d|/max|ref| {rel:.3e}")
    assert cos.min() >= 0.999 and rel <= 2e-2
    sd = ohead.make_head_state(384, 9, 128, 64, seed=8, scale=2.0)
    head = ClassifierLSTMDeltas(384, 9, seq_len=31)
    head.load_state_dict(sd)
    csv = cbas.infer_file(out, head, "m", BEHAVIORS, 31, device=torch.device("cuda"))
    import pandas as pd
    got = pd.read_csv(csv).to_numpy()
    want_p = ohead.infer_windows(emb, sd, seq_len=31)
    assert np.abs(got - want_p).max() <= 1e-3 and (got.argmax(1) == want_p.argmax(1)).all()


def test_worker_threads_encode_then_classify(tmp_path, monkeypatch):
    frames = oenc.synthetic_frames(40, 64, 64, seed=12)
    clips = []
    for i in range(2):
        p = str(tmp_path / f"cam_{i:05d}.npy")
        np.save(p, frames[i * 20:(i + 1) * 20])
        clips.append(p)
    sd = ohead.make_head_state(384, 4, 128, 64, seed=6)
    head = ClassifierLSTMDeltas(384, 4, seq_len=31)
    head.load_state_dict(sd)
    mdir = str(tmp_path / "models" / "M")
    bundle.save_model_bundle(mdir, head, "M", ["a", "b", "c", "d"], 31, "synthetic:vits16", 1.0)
    monkeypatch.setattr(gui_state, "proj", types.SimpleNamespace(encoder_model_identifier="synthetic:vits16"))
    monkeypatch.setattr(gui_state, "dino_encoder", DinoEncoder("synthetic:vits16", "cuda", max_frames=32))
    monkeypatch.setattr(gui_state, "live_inference_model_name", "M")
    gui_state.encode_tasks[:] = clips + [str(tmp_path / "missing.mp4")]
    gui_state.classify_tasks[:] = []
    enc_t = workthreads.EncodeThread("cuda", poll_seconds=0.01)
    cls_t = workthreads.ClassificationThread("cuda", {"M": mdir}, poll_seconds=0.01)
    enc_t.start()
    cls_t.start()
    want = [p.replace(".npy", "_M_outputs.csv") for p in clips]
    for _ in range(3000):
        if all(os.path.exists(w) for w in want):
            break
        time.sleep(0.01)
    enc_t.stop()
    cls_t.stop()
    enc_t.join(5)
    cls_t.join(5)
    assert all(os.path.exists(w) for w in want)           # the bad path was logged and skipped
    import pandas as pd
    for w in want:
        df = pd.read_csv(w)
        assert list(df.columns) == ["a", "b", "c", "d"] and len(df) == 20
        np.testing.assert_allclose(df.to_numpy().sum(1), 1.0, atol=1e-5)


def test_two_host_threads_on_their_own_streams():
    """SURVEY 8b threading contract: EncodeThread and ClassificationThread call into the library concurrently, each
    inside `with torch.cuda.stream(own_stream)`.  Both must get exactly the results they get alone."""
    import threading
    enc = DinoEncoder("synthetic:vits16@2", "cuda", max_frames=16)
    frames = torch.from_numpy(oenc.synthetic_frames(16, 224, 224, seed=41)).cuda()
    sd = ohead.make_head_state(384, 9, 128, 64, seed=9, scale=2.0)
    head = ClassifierLSTMDeltas(384, 9, seq_len=31)
    head.load_state_dict(sd)
    head = head.to("cuda").eval()
    emb = torch.randn(5000, 384, device="cuda").half()
    want_e = enc.encode_u8(frames).clone()
    want_p = head.infer_embeddings(emb).clone()
    torch.cuda.synchronize()
    errors = []

    def encode_loop():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(25):
                    got = enc.encode_u8(frames)
                    s.synchronize()
                    assert torch.equal(got, want_e)
        except Exception as e:  # noqa: BLE001
            errors.append(("encode", repr(e)))

    def classify_loop():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(25):
                    got = head.infer_embeddings(emb)
                    s.synchronize()
                    assert torch.equal(got, want_p)
        except Exception as e:  # noqa: BLE001
            errors.append(("classify", repr(e)))

    ts = [threading.Thread(target=encode_loop), threading.Thread(target=classify_loop)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a worker thread hung"
    assert not errors, errors


def test_launcher_two_ranks_share_the_work(tmp_path):
    """cbas_b200.launch under torchrun with two ranks (both on GPU 0, gloo for the one collective): the videos are
    split between the ranks, every video ends up encoded and classified exactly once, the reduced actogram equals
    the single-process one, and a second run finds nothing left to do (file-level resume)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vids = tmp_path / "rec"
    vids.mkdir()
    for i, n in enumerate((40, 25, 33, 18, 29)):
        np.save(str(vids / f"cam{i}_00001.npy"), oenc.synthetic_frames(n, 64, 64, seed=50 + i))
    sd = ohead.make_head_state(384, 4, 128, 64, seed=6)
    head = ClassifierLSTMDeltas(384, 4, seq_len=31)
    head.load_state_dict(sd)
    mdir = str(tmp_path / "models" / "M")
    bundle.save_model_bundle(mdir, head, "M", ["a", "b", "c", "d"], 31, encoder_model_identifier="synthetic:vits16@4")
    common = ["--videos", str(vids / "*.npy"), "--encoder", "synthetic:vits16@4", "--model-dir", mdir,
              "--actogram", "b", "--framerate", "0.05", "--bin-minutes", "3", "--threshold", "0.0"]
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=root, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])

    two = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29571", "-m", "cbas_b200.launch", *common, "--backend", "gloo"])
    assert two["world_size"] == 2 and two["frames"] == 145
    assert sum(r["encoded"] for r in two["per_rank"]) == 5 and all(r["encoded"] > 0 for r in two["per_rank"])
    assert sum(r["classified"] for r in two["per_rank"]) == 5
    again = run([sys.executable, "-m", "cbas_b200.launch", *common])  # single process, everything already on disk
    assert again["per_rank"][0]["encoded"] == 0 and again["per_rank"][0]["classified"] == 0
    assert again["actogram"]["bins"] == two["actogram"]["bins"] and sum(two["actogram"]["bins"]) > 0


def test_launcher_splits_one_long_video_across_two_ranks(tmp_path):
    """cbas_b200.launch --split-video under torchrun with two ranks (both on GPU 0, gloo): each rank encodes half of
    the frames, rank 0 concatenates the parts, each rank classifies its span with +-15 frames of context.  The
    `_cls.h5` must be bit-identical to the single-process file and the CSV equal to the single-process CSV, cut
    included; nothing but the final artefacts is left on disk."""
    import json
    import subprocess
    import sys
    import pandas as pd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    frames = oenc.synthetic_frames(75, 64, 64, seed=77)
    sd = ohead.make_head_state(384, 4, 128, 64, seed=6)
    head = ClassifierLSTMDeltas(384, 4, seq_len=31)
    head.load_state_dict(sd)
    mdir = str(tmp_path / "models" / "M")
    bundle.save_model_bundle(mdir, head, "M", ["a", "b", "c", "d"], 31, encoder_model_identifier="synthetic:vits16@4")
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    outs = {}
    for tag, pre in (("one", [sys.executable]),
                     ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29573"])):
        d = tmp_path / tag
        d.mkdir()
        np.save(str(d / "day_00001.npy"), frames)
        cmd = pre + ["-m", "cbas_b200.launch", "--videos", str(d / "*.npy"), "--encoder", "synthetic:vits16@4",
                     "--model-dir", mdir, "--actogram", "b", "--framerate", "0.05", "--bin-minutes", "3", "--threshold", "0.0"]
        if tag == "two":
            cmd += ["--backend", "gloo", "--split-video"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=root, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        rep = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        with store.EmbeddingReader(str(d / "day_00001_cls.h5")) as rd:
            emb = rd.read(0, rd.shape[0])
        outs[tag] = (rep, emb, pd.read_csv(str(d / "day_00001_M_outputs.csv")).to_numpy(), sorted(os.listdir(d)))
    one, two = outs["one"], outs["two"]
    assert two[0]["world_size"] == 2 and two[0]["frames"] == 75
    assert [r["frames"] for r in two[0]["per_rank"]] == [38, 37]
    assert two[1].shape == (75, 384) and np.array_equal(one[1], two[1])
    assert np.abs(one[2] - two[2]).max() <= 1e-6
    assert one[3] == two[3], two[3]  # no part files, no .tmp
    assert one[0]["actogram"]["bins"] == two[0]["actogram"]["bins"]


@pytest.mark.parametrize("chunk", [100, 1000, 4096])
def test_infer_file_streams_the_embedding_file_in_chunks(tmp_path, monkeypatch, chunk):
    """infer_file reads INFERENCE_CHUNK_SIZE target frames at a time with +-seq_len//2 rows of context (the reference's
    bounded-memory loop, cbas.py:482,497-508): chunk boundaries - including chunks shorter than the window and a last
    chunk of one frame - must not show in the CSV, which is checked against one whole-file pass and against the oracle
    on a sample of frames."""
    import pandas as pd
    n = 4097
    rng = np.random.default_rng(7)
    emb = (np.cumsum(rng.standard_normal((n, 384)).astype(np.float32) * 0.2, axis=0) % 4.0 - 2.0).astype(np.float16)
    path = str(tmp_path / "long_cls.h5")
    w = store.EmbeddingWriter(path, 384, {})
    w.append(emb.astype(np.float32))
    w.close()
    sd = ohead.make_head_state(384, 9, 128, 64, seed=11, scale=2.0)
    head = ClassifierLSTMDeltas(384, 9, seq_len=31)
    head.load_state_dict(sd)
    whole = pd.read_csv(cbas.infer_file(path, head, "whole", BEHAVIORS, 31, device=torch.device("cuda"))).to_numpy()
    monkeypatch.setattr(cbas, "INFERENCE_CHUNK_SIZE", chunk)
    got = pd.read_csv(cbas.infer_file(path, head, "chunked", BEHAVIORS, 31, device=torch.device("cuda"))).to_numpy()
    assert got.shape == whole.shape == (n, 9)
    assert np.abs(got - whole).max() <= 1e-6 and (got.argmax(1) == whole.argmax(1)).all()
    sample = np.r_[0:40, chunk - 20:chunk + 20, n - 40:n]
    sample = sample[(sample >= 0) & (sample < n)]
    want = ohead.infer_windows(emb, sd, seq_len=31)[sample]
    assert np.abs(got[sample] - want).max() <= 1e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(tmp_path, monkeypatch):
    """One process, two GPUs (workthreads.py: a thread pair per device): an encoder and a head created on cuda:1 are
    driven from threads whose current device is cuda:0 - every native handle makes its own device current per call -
    while a second pair works on cuda:0; results equal the single-device ones, and the per-device kernel attributes
    (shared-memory opt-in, SM count) are set up on both devices."""
    frames = oenc.synthetic_frames(24, 224, 224, seed=21)
    sd = ohead.make_head_state(384, 4, 128, 64, seed=6)
    encs, heads, outs = {}, {}, {}
    for d in (0, 1):
        encs[d] = DinoEncoder("synthetic:vits16@4", f"cuda:{d}", max_frames=24)
        heads[d] = ClassifierLSTMDeltas(384, 4, seq_len=31)
        heads[d].load_state_dict(sd)
        heads[d].to(f"cuda:{d}")
    torch.cuda.set_device(0)
    errors = []

    def work(d):
        try:
            assert torch.cuda.current_device() == 0  # threads inherit device 0: the handles must switch themselves
            emb = encs[d].encode_u8(torch.from_numpy(frames).to(f"cuda:{d}"))
            probs = heads[d].infer_embeddings(emb.half())
            from cbas_b200.classifier_head import actogram_bins
            outs[d] = (emb.cpu(), probs.cpu(), actogram_bins(probs, 1, 0.2, 5).cpu())
        except Exception as e:  # noqa: BLE001
            errors.append((d, repr(e)))

    ts = [threading.Thread(target=work, args=(d,)) for d in (1, 0)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(300)
    assert not errors, errors
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    # EncodeThread pair per device over one shared queue
    clips = []
    for i in range(4):
        p = str(tmp_path / f"cam_{i:05d}.npy")
        np.save(p, oenc.synthetic_frames(12, 64, 64, seed=60 + i))
        clips.append(p)
    monkeypatch.setattr(gui_state, "proj", None)
    monkeypatch.setattr(gui_state, "dino_encoder", None)
    monkeypatch.setattr(gui_state, "dino_encoders", {"cuda:0": encs[0], "cuda:1": encs[1]})
    monkeypatch.setattr(gui_state, "encode_tasks", list(clips))
    monkeypatch.setattr(gui_state, "live_inference_model_name", None)
    workers = [workthreads.EncodeThread("cuda:0", poll_seconds=0.02), workthreads.EncodeThread("cuda:1", poll_seconds=0.02)]
    for w in workers:
        w.start()
    deadline = time.time() + 120
    while time.time() < deadline and not all(os.path.exists(c[:-4] + "_cls.h5") for c in clips):
        time.sleep(0.05)
    for w in workers:
        w.stop()
    for w in workers:
        w.join(10)
    assert all(os.path.exists(c[:-4] + "_cls.h5") for c in clips)
    assert sum(w.tasks_processed_in_batch for w in workers) >= 0


def test_encode_file_parallel_decode_threads_are_bit_identical(tmp_path, monkeypatch):
    """mp4 -> `_cls.h5` with one in-process decoder (the reference's arrangement) and with four decoding whole chunks
    side by side: the same embedding file, bit for bit, in both preprocessing modes."""
    cv2 = pytest.importorskip("cv2")
    p = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (64, 64))
    if not vw.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    base = np.random.default_rng(3).integers(0, 255, (64, 64, 3), dtype=np.uint8)
    for i in range(230):
        vw.write(np.roll(base, 3 * i, axis=1))
    vw.release()
    monkeypatch.setattr(gui_state, "proj", None)
    monkeypatch.setattr(cbas, "CHUNK_SIZE", 32)
    monkeypatch.setattr(cbas, "DECODE_WORKERS", 0)
    for mode in ("reference", "processor"):
        enc = DinoEncoder("synthetic:vits16@4", "cuda", preprocess=mode, image_size=64, max_frames=32)
        outs = []
        for threads in (1, 4):
            monkeypatch.setattr(cbas, "DECODE_THREADS", threads)
            out = cbas.encode_file(enc, p)
            with store.EmbeddingReader(out) as r:
                outs.append(r.read(0, r.shape[0]))
            os.remove(out)
        assert outs[0].shape == (230, 384) and np.array_equal(outs[0], outs[1]), mode
