"""Host-layer logic that runs without a GPU: dataset formats of the reference drivers (reference filehelper.cpp:106-345), through the C++ host layer.
CPU-only: the files are written by numpy in the documented layout and read back through the library, and vice versa."""
import numpy as np
import pytest

from vector_line_quantization_b200 import filehelper as fh
from vector_line_quantization_b200.index import FaissException


def _texmex_bytes(x):
    n, d = x.shape
    rec = np.empty((n, 4 + d * x.dtype.itemsize), dtype=np.uint8)
    rec[:, :4] = np.frombuffer(np.int32(d).tobytes(), dtype=np.uint8)
    rec[:, 4:] = x.view(np.uint8).reshape(n, -1)
    return rec.tobytes()


@pytest.mark.parametrize("kind,dt", [("fvecs", np.float32), ("ivecs", np.int32), ("bvecs", np.uint8)])
def test_texmex_read_matches_layout(tmp_path, kind, dt):
    rng = np.random.default_rng(0)
    x = (rng.random((37, 24)) * 200).astype(dt)
    p = tmp_path / ("base." + kind)
    p.write_bytes(_texmex_bytes(x))
    assert fh.readJegouHeader(p) == (37, 24)
    np.testing.assert_array_equal(fh.readJegou(p), x)
    np.testing.assert_array_equal(fh.readJegou(p, start=5, num=9), x[5:14])   # readBatchJegou
    np.testing.assert_array_equal(fh.readJegou(p, start=30, num=100), x[30:])  # clipped at the end
    assert fh.readJegou(p, start=37).shape == (0, 24)


@pytest.mark.parametrize("kind,dt", [("fvecs", np.float32), ("ivecs", np.int32), ("bvecs", np.uint8)])
def test_texmex_write_is_byte_exact(tmp_path, kind, dt):
    x = (np.arange(5 * 8).reshape(5, 8) % 251).astype(dt)
    p = tmp_path / ("w." + kind)
    fh.writeJegou(p, x)
    assert p.read_bytes() == _texmex_bytes(x)


def test_texmex_errors(tmp_path):
    with pytest.raises(FaissException):
        fh.readJegou(tmp_path / "missing.fvecs")
    p = tmp_path / "trunc.fvecs"
    p.write_bytes(_texmex_bytes(np.ones((3, 4), np.float32))[:-3])
    with pytest.raises(FaissException):
        fh.readJegou(p)


def test_umem_layout_and_offsets(tmp_path):
    rng = np.random.default_rng(1)
    num, dim = 11, 6
    x = rng.random((num, dim)).astype(np.float32)
    p = tmp_path / "m.umem"
    fh.write(p, num, dim, x[:4], offset=0)            # creates header + first rows
    fh.write(p, num, dim, x[4:], offset=4 * dim)      # appended in place, like the reference's chunked writers
    raw = p.read_bytes()
    assert raw[:6] == b"11\n6\n" + b"\0"               # ASCII header, padded to the fixed payload offset
    assert len(raw) == fh.UMEM_PAYLOAD_OFFSET + x.nbytes
    np.testing.assert_array_equal(np.frombuffer(raw[20:], dtype=np.float32).reshape(num, dim), x)
    assert fh.header(p) == (num, dim)
    np.testing.assert_array_equal(fh.readFloat(p, dim, num), x)
    np.testing.assert_array_equal(fh.readFloat(p, dim, 3, offset=5), x[5:8])


def test_imem_and_u8(tmp_path):
    ids = np.arange(40, dtype=np.int32).reshape(10, 4) * 7
    p = tmp_path / "gt.imem"
    fh.write(p, 10, 4, ids)
    np.testing.assert_array_equal(fh.readInt(p, 4, 10), ids)
    codes = (np.arange(64) % 256).astype(np.uint8).reshape(4, 16)
    q = tmp_path / "c.umem"
    fh.write(q, 4, 16, codes)
    np.testing.assert_array_equal(fh.readUint8(q, 16, 2, offset=1), codes[1:3])
    with pytest.raises(FaissException):
        fh.readInt(p, 4, 11)                           # short read


# ------------------------------------------------------------------------------------------------ host k-means RNG
@pytest.mark.parametrize("n,seed", [(1, 1234), (10, 1234), (1000, 1235), (65536, 7)])
def test_host_rand_perm_matches_reference(n, seed):
    """host/Clustering.cpp rand_perm (the subsampling / initialisation order of the device k-means) against the
    reference library's rand_perm (utils.cpp:307-317) and the oracle restatement"""
    import ctypes as C

    from oracle import pyoracle as po
    from vector_line_quantization_b200.index import _call

    perm = np.empty(n, dtype=np.int32)
    _call("vlq_host_rand_perm", C.c_void_p(perm.ctypes.data), C.c_long(n), C.c_long(seed))
    assert sorted(perm.tolist()) == list(range(n))
    np.testing.assert_array_equal(perm, po.rand_perm(n, seed))
    if po.ref_available():
        np.testing.assert_array_equal(perm, po.ref_rand_perm(n, seed))
