"""Property-based tests (hypothesis) of the host-side pieces around the kernels: the native HDF5 writer/reader round
trip, the event run-length encoding against a plain state machine written from the reference's description, and the
multi-GPU work partitioning."""
import os

import numpy as np
from hypothesis import given, settings, strategies as st

from cbas_b200 import events, hdf5_min, parallel

SET = settings(max_examples=40, deadline=None)


@SET
@given(rows=st.integers(0, 700), width=st.sampled_from([1, 3, 64, 384]), chunk=st.sampled_from([1, 7, 64, 8192]),
       pieces=st.integers(1, 5), seed=st.integers(0, 2 ** 16),
       attr=st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=0x2FF), min_size=0, max_size=40))
def test_hdf5_min_round_trip(tmp_path_factory, rows, width, chunk, pieces, seed, attr):
    """Any number of rows appended in any split, any chunk length, float16 payload and UTF-8 attributes come back
    exactly through the independent reader (which is itself validated against a file written by the HDF5 library)."""
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((rows, width)).astype(np.float16)
    path = str(tmp_path_factory.mktemp("h5") / "x_cls.h5")
    w = hdf5_min.Writer(path, "cls", width, "f2", chunk_rows=chunk,
                        attrs={"encoder_model_identifier": attr, "schema_version": "1.0"})
    cuts = sorted(rng.integers(0, rows + 1, size=pieces - 1).tolist()) if rows else []
    for a, b in zip([0] + cuts, cuts + [rows]):
        w.append(data[a:b])
    w.close()
    with hdf5_min.File(path) as f:
        ds = f["cls"]
        assert ds.shape == (rows, width) and ds.dtype == np.float16
        assert np.array_equal(ds[:], data)
        if rows:
            a, b = sorted(rng.integers(0, rows + 1, size=2).tolist())
            assert np.array_equal(ds.read_rows(a, b), data[a:b])
        assert f.attrs["encoder_model_identifier"] == attr and f.attrs["schema_version"] == "1.0"
    os.remove(path)


def _instances_state_machine(p, behaviors, thr):
    """The reference's description, as a per-frame state machine (cbas.py:903-929 in words): an event opens on a
    frame whose top probability reaches the threshold, closes before the first frame that is below it or has another
    top behaviour, and one that is still open at the end closes on the last frame."""
    out, cur = [], None
    for i, row in enumerate(p):
        lab = int(np.argmax(row))  # first maximum
        ok = row[lab] >= thr
        if cur is not None and (not ok or lab != cur[1]):
            out.append((cur[0], i - 1, behaviors[cur[1]]))
            cur = None
        if cur is None and ok:
            cur = (i, lab)
    if cur is not None:
        out.append((cur[0], len(p) - 1, behaviors[cur[1]]))
    return out


@SET
@given(n=st.integers(0, 300), C=st.integers(1, 6), thr=st.floats(0.0, 1.0), seed=st.integers(0, 2 ** 16),
       runs=st.booleans())
def test_event_extraction_matches_state_machine(n, C, thr, seed, runs):
    rng = np.random.default_rng(seed)
    p = rng.random((n, C))
    if runs and n:  # long constant stretches and exact ties
        p = np.repeat(p[:: max(1, n // 7 + 1)], n // 7 + 1, axis=0)[:n]
        p = np.round(p, 1)
    behaviors = [f"b{i}" for i in range(C)]
    got = [(d["start"], d["end"], d["label"]) for d in events.predictions_to_instances(p, "m", behaviors, thr, video="v")]
    assert got == _instances_state_machine(p, behaviors, thr)
    blocks, _ = events.predictions_to_instances_with_confidence(p, "m", behaviors, video="v")
    # blocks tile the whole clip without gaps or overlaps and carry the mean top probability
    if n:
        assert blocks[0]["start"] == 0 and blocks[-1]["end"] == n - 1
        for a, b in zip(blocks, blocks[1:]):
            assert b["start"] == a["end"] + 1 and a["label"] != b["label"]
        top = p.max(axis=1)
        for b in blocks:
            assert abs(b["confidence"] - top[b["start"]:b["end"] + 1].mean()) < 1e-9
    else:
        assert blocks == []


@SET
@given(costs=st.lists(st.floats(0.1, 1e4), min_size=0, max_size=60), world=st.integers(1, 8))
def test_partition_videos_is_a_balanced_partition(costs, world):
    paths = [f"cam{i % 5}/seg{i:03d}.mp4" for i in range(len(costs))]
    shards = parallel.partition_videos(paths, costs, world)
    assert len(shards) == world and sorted(sum(shards, [])) == sorted(paths)      # every video exactly once
    assert shards == parallel.partition_videos(list(reversed(paths)), list(reversed(costs)), world)  # order-free
    if costs:
        cost = dict(zip(paths, costs))
        loads = [sum(cost[p] for p in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs) + 1e-6                       # LPT: within one item of each other


@SET
@given(n=st.integers(0, 100000), world=st.integers(1, 8), halo=st.integers(0, 47))
def test_split_frame_range_covers_every_frame_once(n, world, halo):
    spans = parallel.split_frame_range(n, world, halo)
    core = parallel.split_frame_range(n, world, 0)
    assert sum(len(r) for r in core) == n and all(a.stop == b.start for a, b in zip(core, core[1:]))
    for s, c in zip(spans, core):
        assert s.start == max(0, c.start - halo) and s.stop == min(n, c.stop + halo)
