"""Dataset / matrix file formats of the reference drivers, read and written by the C++ host layer
(host/filehelper.{h,cpp}; reference filehelper.cpp:106-345): TexMex .fvecs/.ivecs/.bvecs and .umem/.imem
("num\\ndim\\n" ASCII header, payload at byte 20).  Thin ctypes plumbing; names follow the reference's helpers."""
import ctypes as C

import numpy as np

from .index import _call

_KINDS = {"fvecs": (0, np.float32), "ivecs": (1, np.int32), "bvecs": (2, np.uint8)}
UMEM_PAYLOAD_OFFSET = 20


def _kind(path, kind):
    kind = kind or str(path).rsplit(".", 1)[-1]
    if kind not in _KINDS:
        raise ValueError("unknown TexMex kind %r" % kind)
    return _KINDS[kind]


def readJegouHeader(path, kind=None):
    """-> (n, d) of a TexMex file (reference readJegouHeader)."""
    _, dt = _kind(path, kind)
    n, d = C.c_long(), C.c_long()
    _call("vlq_host_vecs_header", str(path).encode(), np.dtype(dt).itemsize, C.byref(n), C.byref(d))
    return n.value, d.value


def readJegou(path, start=0, num=0, kind=None):
    """Vectors [start, start+num) (num=0: to the end) as an (n, d) array (reference readJegou / readBatchJegou)."""
    code, dt = _kind(path, kind)
    n, d = readJegouHeader(path, kind)
    if start > n:
        raise ValueError("start beyond the end of the file")
    cnt = n - start if num == 0 else min(num, n - start)
    out = np.empty((cnt, d), dtype=dt)
    _call("vlq_host_vecs_read", str(path).encode(), code, C.c_long(start), C.c_long(cnt), C.c_void_p(out.ctypes.data))
    return out


def writeJegou(path, x, kind=None):
    code, dt = _kind(path, kind)
    x = np.ascontiguousarray(x, dtype=dt)
    _call("vlq_host_vecs_write", str(path).encode(), code, C.c_void_p(x.ctypes.data), C.c_long(x.shape[0]),
          C.c_long(x.shape[1]))


def header(path):
    """-> (num, dim) of a .umem/.imem file (reference header())."""
    n, d = C.c_long(), C.c_long()
    _call("vlq_host_umem_header", str(path).encode(), C.byref(n), C.byref(d))
    return n.value, d.value


def write(path, num, dim, x, offset=0):
    """Write the flat array x at element offset `offset` of the payload; offset 0 creates the file (reference write<T>)."""
    x = np.ascontiguousarray(x)
    _call("vlq_host_umem_write", str(path).encode(), C.c_long(num), C.c_long(dim), C.c_void_p(x.ctypes.data),
          x.dtype.itemsize, C.c_long(x.size), C.c_long(offset))


def read(path, dtype, length, offset=0):
    """Read `length` elements of `dtype` from element offset `offset` (reference read<T> / readFloat / readInt)."""
    out = np.empty(length, dtype=dtype)
    _call("vlq_host_umem_read", str(path).encode(), C.c_void_p(out.ctypes.data), out.dtype.itemsize, C.c_long(length),
          C.c_long(offset))
    return out


def readFloat(path, dim, num, offset=0):
    return read(path, np.float32, dim * num, offset * dim).reshape(num, dim)


def readUint8(path, dim, num, offset=0):
    return read(path, np.uint8, dim * num, offset * dim).reshape(num, dim)


def readInt(path, dim, num, offset=0):
    return read(path, np.int32, dim * num, offset * dim).reshape(num, dim)
