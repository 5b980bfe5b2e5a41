"""CPU: the native HDF5 implementation - reader against a genuine libhdf5-written file, writer by round trip and
by comparing its framing with that file; the EmbeddingStore contract of cbas.py:413-421."""
import glob
import os
import struct

import numpy as np
import pytest

from cbas_b200 import hdf5_min, store

REAL = glob.glob(os.path.join(os.path.dirname(__import__("scipy").__file__), "io/matlab/tests/data/testhdf5_7.4_GLNX86.mat"))


@pytest.mark.skipif(not REAL, reason="scipy's HDF5 sample file not present")
def test_reader_on_genuine_libhdf5_file():
    # MATLAB v7.3 = HDF5 1.6 behind a 512-byte user block: superblock search, base address, group B-tree, SNOD,
    # local heap, v1 object header, old layout message, fixed-length string attribute
    import scipy.io
    with hdf5_min.File(REAL[0]) as f:
        assert list(f.keys()) == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        want = scipy.io.loadmat(REAL[0].replace("testhdf5", "testdouble"))["testdouble"].ravel()
        np.testing.assert_array_equal(d[...].ravel(), want)
        assert f._rd.attributes(f._links["testdouble"])["MATLAB_class"] == b"double"


@pytest.mark.parametrize("n", [0, 1, 511, 8192, 8193, 20000])
def test_writer_round_trip(tmp_path, n):
    rng = np.random.default_rng(n)
    emb = rng.standard_normal((n, 768)).astype(np.float32)
    p = str(tmp_path / "clip_cls.h5")
    w = store.EmbeddingWriter(p, 768, {"encoder_model_identifier": "facebook/dinov3-vitb16-pretrain-lvd1689m",
                                       "schema_version": "1.0"}, backend="native")
    for i in range(0, n, 512):  # the reference appends 512-frame chunks (cbas.py:423-440)
        w.append(emb[i:i + 512])
        w.flush()
    w.close()
    with store.EmbeddingReader(p, backend="native") as r:
        assert r.shape == (n, 768)
        assert r.attrs == {"encoder_model_identifier": "facebook/dinov3-vitb16-pretrain-lvd1689m", "schema_version": "1.0"}
        assert isinstance(r.attrs["encoder_model_identifier"], str)  # compared with a str in startup_page.py:110
        np.testing.assert_array_equal(r.read(0, n), emb.astype(np.float16))
        if n > 9000:  # slices across the chunk boundary, as infer_file reads them (cbas.py:503-507)
            np.testing.assert_array_equal(r.read(8185, 8215), emb[8185:8215].astype(np.float16))
            np.testing.assert_array_equal(r.read(n - 7, n + 100), emb[n - 7:].astype(np.float16))
    with hdf5_min.File(p) as f:
        d = f["cls"]
        assert d.dtype == np.dtype("<f2") and d.chunks == (8192, 768) and d.maxshape == (None, 768)
        assert len(d) == n


def test_writer_framing_matches_libhdf5_conventions(tmp_path):
    p = str(tmp_path / "x_cls.h5")
    w = hdf5_min.Writer(p, "cls", 384, "f2", 8192, {"schema_version": "1.0"})
    w.append(np.ones((3, 384), np.float16))
    w.close()
    raw = open(p, "rb").read()
    assert raw[:8] == hdf5_min.SIG
    ver, _, _, _, _, so, sl, _, leaf_k, int_k = struct.unpack_from("<BBBBBBBBHH", raw, 8)
    assert (ver, so, sl, leaf_k, int_k) == (0, 8, 8, 4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == hdf5_min.UNDEF and drv == hdf5_min.UNDEF and eof == len(raw)
    root = struct.unpack_from("<Q", raw, 56 + 8)[0]
    assert raw[root] == 1 and root % 8 == 0          # v1 object header, 8-byte aligned
    btree, heap = struct.unpack_from("<QQ", raw, 56 + 24)
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    if REAL:  # same 16-byte object-header prefix shape and the same group-node / heap conventions as libhdf5
        real = open(REAL[0], "rb").read()[512:]
        rroot = struct.unpack_from("<Q", real, 56 + 8)[0]
        assert real[rroot] == raw[root] == 1 and real[rroot + 12:rroot + 16] == raw[root + 12:root + 16] == b"\0" * 4
        rbt, rheap = struct.unpack_from("<QQ", real, 56 + 24)
        assert real[rbt:rbt + 8] == raw[btree:btree + 8]              # TREE, type 0, level 0, 1 entry
        assert struct.unpack_from("<Q", real, rbt + 24)[0] == struct.unpack_from("<Q", raw, btree + 24)[0] == 0
        assert real[rheap:rheap + 8] == raw[heap:heap + 8]            # HEAP, version 0
        rfree = struct.unpack_from("<Q", real, rheap + 16)[0]
        rdata = struct.unpack_from("<Q", real, rheap + 24)[0]
        mfree = struct.unpack_from("<Q", raw, heap + 16)[0]
        mdata = struct.unpack_from("<Q", raw, heap + 24)[0]
        assert struct.unpack_from("<Q", real, rdata + rfree)[0] == struct.unpack_from("<Q", raw, mdata + mfree)[0] == 1


def test_corrupt_and_foreign_files(tmp_path):
    p = tmp_path / "bad_cls.h5"
    p.write_bytes(b"not hdf5 at all" * 100)
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.File(str(p))
    w = hdf5_min.Writer(str(tmp_path / "t.h5"), "other", 8, "f2")
    w.close()
    with pytest.raises(KeyError):
        store.EmbeddingReader(str(tmp_path / "t.h5"), backend="native")


def test_interop_with_h5py_both_directions(tmp_path):
    """Runs only where h5py (the reference's own HDF5 library) is importable - it is in neither development image: a
    file from the native writer opened by libhdf5 (shape, dtype, chunking, unlimited first axis, string attributes,
    data), and a file h5py wrote read by the native reader and through store.EmbeddingReader."""
    h5py = pytest.importorskip("h5py")
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((9000, 768)).astype(np.float32)
    attrs = {"encoder_model_identifier": "facebook/dinov3-vitb16-pretrain-lvd1689m", "schema_version": "1.0"}
    p = str(tmp_path / "native_cls.h5")
    w = store.EmbeddingWriter(p, 768, attrs, backend="native")
    for i in range(0, 9000, 512):
        w.append(emb[i:i + 512])
        w.flush()
    w.close()
    with h5py.File(p, "r") as f:
        d = f["cls"]
        assert d.shape == (9000, 768) and d.dtype == np.float16 and d.maxshape == (None, 768)
        assert d.chunks == (store.CHUNK_ROWS, 768)
        assert {k: (v.decode() if isinstance(v, bytes) else str(v)) for k, v in f.attrs.items()} == attrs
        np.testing.assert_array_equal(d[...], emb.astype(np.float16))
        np.testing.assert_array_equal(d[4000:4100], emb[4000:4100].astype(np.float16))
    q = str(tmp_path / "h5py_cls.h5")
    w = store.EmbeddingWriter(q, 768, attrs, backend="h5py")
    w.append(emb)
    w.close()
    with hdf5_min.File(q) as f:
        assert f["cls"].shape == (9000, 768)
        np.testing.assert_array_equal(f["cls"][100:700], emb[100:700].astype(np.float16))
    with store.EmbeddingReader(q) as r:
        assert r.shape == (9000, 768) and r.attrs == attrs
