"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/libvlq_oracle.so (CPU restatement, oracle/vlq_oracle.h)
and oracle/_ref/libfaiss_ref.so (the unmodified reference CPU library + oracle/ref_shim.cpp).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32
_i32 = np.int32
_i64 = np.int64
_u8 = np.uint8

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)
c_u8_p = C.POINTER(C.c_uint8)


def _p(a, ty):
    if a is None:
        return None
    return a.ctypes.data_as(ty)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def build(verbose=False):
    """(Re)build the oracle and -- when /root/reference exists -- the reference library."""
    out = subprocess.run(["make", "-s", "-C", HERE, "all", "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libvlq_oracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.vlqo_decode_distance.restype = C.c_double
    return _lib


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libfaiss_ref.so"))


def ref():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libfaiss_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libfaiss_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        # OpenBLAS (bundled in the opencv wheel) needs its sibling libgfortran/libquadmath: preload them
        import glob
        blas_dir = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*"):
            for dep in glob.glob(os.path.join(blas_dir, pat)):
                C.CDLL(dep, mode=C.RTLD_GLOBAL)
        _ref = C.CDLL(path)
        _ref.ref_ivfpq_new.restype = C.c_void_p
        _ref.ref_ivfpq_list_size.restype = C.c_long
    return _ref


# --------------------------------------------------------------------------------------------- oracle (vlq_oracle.h)
def num_threads():
    return lib().vlqo_num_threads()


def rand_perm(n, seed):
    perm = np.empty(n, _i32)
    lib().vlqo_rand_perm(_p(perm, c_int_p), C.c_long(n), C.c_long(seed))
    return perm


def l2_topk(x, cent, k, add_xnorm=True):
    x = _c(x, _f32)
    cent = _c(cent, _f32)
    n, d = x.shape
    D = np.empty((n, k), _f32)
    I = np.empty((n, k), _i32)
    lib().vlqo_l2_topk(_p(x, c_float_p), C.c_long(n), d, _p(cent, c_float_p), C.c_long(cent.shape[0]), k,
                       int(add_xnorm), _p(D, c_float_p), _p(I, c_int_p))
    return D, I


def kmeans(x, k, niter=10, seed=1234, max_points_per_centroid=256):
    x = _c(x, _f32)
    n, d = x.shape
    cent = np.empty((k, d), _f32)
    obj = np.empty(niter, _f32)
    lib().vlqo_kmeans(d, k, C.c_long(n), _p(x, c_float_p), niter, C.c_long(seed), max_points_per_centroid,
                      _p(cent, c_float_p), _p(obj, c_float_p))
    return cent, obj


def knn_graph(cent, E):
    cent = _c(cent, _f32)
    Cn, d = cent.shape
    edge = np.empty((Cn, E), _i32)
    ed2 = np.empty((Cn, E), _f32)
    lib().vlqo_knn_graph(_p(cent, c_float_p), C.c_long(Cn), d, E, _p(edge, c_int_p), _p(ed2, c_float_p))
    return edge, ed2


def line_stage(x, A, cent, edge, edge_d2):
    x = _c(x, _f32)
    A = _c(A, _i32)
    n, d = x.shape
    E = edge.shape[1]
    out_list = np.empty(n, _i32)
    out_lam = np.empty(n, _f32)
    lib().vlqo_line_stage(_p(x, c_float_p), C.c_long(n), d, _p(A, c_int_p), _p(_c(cent, _f32), c_float_p),
                          _p(_c(edge, _i32), c_int_p), _p(_c(edge_d2, _f32), c_float_p), E, _p(out_list, c_int_p),
                          _p(out_lam, c_float_p))
    return out_list, out_lam


def lambda_quantize(lam, cb):
    lam = _c(lam, _f32)
    cb = _c(cb, _f32)
    out = np.empty(lam.shape[0], _u8)
    lib().vlqo_lambda_quantize(_p(lam, c_float_p), C.c_long(lam.shape[0]), _p(cb, c_float_p), cb.shape[0],
                               _p(out, c_u8_p))
    return out


def residual(x, list_ids, lamq, cb, cent, edge):
    x = _c(x, _f32)
    n, d = x.shape
    r = np.empty((n, d), _f32)
    lib().vlqo_residual(_p(x, c_float_p), C.c_long(n), d, _p(_c(list_ids, _i32), c_int_p), _p(_c(lamq, _u8), c_u8_p),
                        _p(_c(cb, _f32), c_float_p), _p(_c(cent, _f32), c_float_p), _p(_c(edge, _i32), c_int_p),
                        edge.shape[1], _p(r, c_float_p))
    return r


def pq_encode(r, pq):
    """pq: (M, ksub, dsub)"""
    r = _c(r, _f32)
    pq = _c(pq, _f32)
    n, d = r.shape
    M, ksub, _ = pq.shape
    codes = np.empty((n, M), _u8)
    lib().vlqo_pq_encode(_p(r, c_float_p), C.c_long(n), d, _p(pq, c_float_p), M, ksub, _p(codes, c_u8_p))
    return codes


def build_lists(list_ids, nlists):
    list_ids = _c(list_ids, _i32)
    n = list_ids.shape[0]
    offsets = np.empty(nlists + 1, _i64)
    perm = np.empty(n, _i64)
    lib().vlqo_build_lists(_p(list_ids, c_int_p), C.c_long(n), C.c_long(nlists), _p(offsets, c_long_p),
                           _p(perm, c_long_p))
    return offsets, perm


def term2(cent, pq):
    cent = _c(cent, _f32)
    pq = _c(pq, _f32)
    Cn, d = cent.shape
    M, ksub, _ = pq.shape
    T2 = np.empty((Cn, M, ksub), _f32)
    lib().vlqo_term2(_p(cent, c_float_p), C.c_long(Cn), d, _p(pq, c_float_p), M, ksub, _p(T2, c_float_p))
    return T2


def search(q, cent, edge, edge_d2, lambda_cb, pq, offsets, codes, lamq, ids, P, W, k, cap=1024, T2=None,
           want_debug=False):
    q = _c(q, _f32)
    cent = _c(cent, _f32)
    nq, d = q.shape
    Cn = cent.shape[0]
    E = edge.shape[1]
    M, ksub, _ = pq.shape
    D = np.empty((nq, k), _f32)
    I = np.empty((nq, k), _i64)
    Pe = min(P, Cn)
    coarse = np.empty((nq, Pe), _i32) if want_debug else None
    lines = np.empty((nq, W), _i32) if want_debug else None
    nscan = np.empty(nq, _i64) if want_debug else None
    if T2 is not None:
        T2 = _c(T2, _f32)
    lib().vlqo_search(_p(q, c_float_p), C.c_long(nq), d, _p(cent, c_float_p), C.c_long(Cn),
                      _p(_c(edge, _i32), c_int_p), _p(_c(edge_d2, _f32), c_float_p), E,
                      _p(_c(lambda_cb, _f32), c_float_p), len(lambda_cb), _p(_c(pq, _f32), c_float_p), M, ksub,
                      _p(T2, c_float_p), _p(_c(offsets, _i64), c_long_p), _p(_c(codes, _u8), c_u8_p),
                      _p(_c(lamq, _u8), c_u8_p), _p(_c(ids, _i64), c_long_p), P, W, k, cap, _p(D, c_float_p),
                      _p(I, c_long_p), _p(coarse, c_int_p), _p(lines, c_int_p), _p(nscan, c_long_p))
    if want_debug:
        return D, I, coarse, lines, nscan
    return D, I


def scan_lines(q, cent, edge, edge_d2, lambda_cb, pq, offsets, codes, lamq, ids, lines, k, cap=1024, T2=None):
    """the scan half of search() on a given line choice: lines [nq][W] list ids in rank order (-1 padded)"""
    q = _c(q, _f32)
    cent = _c(cent, _f32)
    lines = _c(lines, _i32)
    nq, d = q.shape
    W = lines.shape[1]
    M, ksub, _ = pq.shape
    D = np.empty((nq, k), _f32)
    I = np.empty((nq, k), _i64)
    if T2 is not None:
        T2 = _c(T2, _f32)
    lib().vlqo_scan_lines(_p(q, c_float_p), C.c_long(nq), d, _p(cent, c_float_p), C.c_long(cent.shape[0]),
                          _p(_c(edge, _i32), c_int_p), _p(_c(edge_d2, _f32), c_float_p), edge.shape[1],
                          _p(_c(lambda_cb, _f32), c_float_p), len(lambda_cb), _p(_c(pq, _f32), c_float_p), M, ksub,
                          _p(T2, c_float_p), _p(_c(offsets, _i64), c_long_p), _p(_c(codes, _u8), c_u8_p),
                          _p(_c(lamq, _u8), c_u8_p), _p(_c(ids, _i64), c_long_p), _p(lines, c_int_p), W, k, cap,
                          _p(D, c_float_p), _p(I, c_long_p))
    return D, I


def merge_topk(D, I):
    """D, I: [R][nq][k]"""
    D = _c(D, _f32)
    I = _c(I, _i64)
    R, nq, k = D.shape
    oD = np.empty((nq, k), _f32)
    oI = np.empty((nq, k), _i64)
    lib().vlqo_merge_topk(_p(D, c_float_p), _p(I, c_long_p), R, C.c_long(nq), k, _p(oD, c_float_p), _p(oI, c_long_p))
    return oD, oI


def imi_search(x, cent, k):
    """IMI coarse quantizer (next row f3): cent = (M, ksub, dsub) sub-space codebooks -> (D [n][k], cell labels [n][k])"""
    x = _c(x, _f32)
    cent = _c(cent, _f32)
    n, d = x.shape
    M, ksub, _ = cent.shape
    D = np.empty((n, k), _f32)
    I = np.empty((n, k), _i64)
    lib().vlqo_imi_search(_p(x, c_float_p), C.c_long(n), d, _p(cent, c_float_p), M, ksub, k, _p(D, c_float_p),
                          _p(I, c_long_p))
    return D, I


def decode_distance(q, c, s, lam, pq, code):
    pq = _c(pq, _f32)
    M, ksub, _ = pq.shape
    return lib().vlqo_decode_distance(_p(_c(q, _f32), c_float_p), q.shape[0], _p(_c(c, _f32), c_float_p),
                                      _p(_c(s, _f32), c_float_p), C.c_float(lam), _p(pq, c_float_p), M, ksub,
                                      _p(_c(code, _u8), c_u8_p))


def encode_all(x, cent, edge, edge_d2, lambda_cb, pq):
    """Full oracle encode (a2,a5-a8): returns dict(A, list, lam, lamq, codes)."""
    _, A = l2_topk(x, cent, 1, add_xnorm=True)
    A = A[:, 0].copy()
    lst, lam = line_stage(x, A, cent, edge, edge_d2)
    lamq = lambda_quantize(lam, lambda_cb)
    r = residual(x, lst, lamq, lambda_cb, cent, edge)
    codes = pq_encode(r, pq)
    return dict(A=A, list=lst, lam=lam, lamq=lamq, codes=codes, residual=r)


def train_all(xt, nlist, E, M, nL, nbits=8, niter=10, pq_niter=25, seed=1234):
    """Oracle restatement of GpuIndexIVFPQ::train (gpu/GpuIndexIVFPQ.cu:1160-1178, 345-403)."""
    xt = _c(xt, _f32)
    d = xt.shape[1]
    cent, _ = kmeans(xt, nlist, niter=niter, seed=seed)
    edge, ed2 = knn_graph(cent, E)
    n2 = min(xt.shape[0], (1 << nbits) * 128)
    x2 = xt[:n2]
    _, A = l2_topk(x2, cent, 1)
    lst, lam = line_stage(x2, A[:, 0].copy(), cent, edge, ed2)
    lcb, _ = kmeans(lam.reshape(-1, 1), nL, niter=niter, seed=seed)
    lcb = lcb.reshape(-1)
    lamq = lambda_quantize(lam, lcb)
    r = residual(x2, lst, lamq, lcb, cent, edge)
    ksub = 1 << nbits
    dsub = d // M
    pq = np.empty((M, ksub, dsub), _f32)
    for m in range(M):
        pq[m], _ = kmeans(r[:, m * dsub:(m + 1) * dsub], ksub, niter=pq_niter, seed=seed)
    return dict(cent=cent, edge=edge, edge_d2=ed2, lambda_cb=lcb, pq=pq)


# -------------------------------------------------------------------------- reference (oracle/_ref/libfaiss_ref.so)
def ref_num_threads():
    return ref().ref_num_threads()


def ref_flat_search(xb, xq, k):
    xb = _c(xb, _f32)
    xq = _c(xq, _f32)
    nq, d = xq.shape
    D = np.empty((nq, k), _f32)
    I = np.empty((nq, k), _i64)
    ref().ref_flat_search(d, C.c_long(xb.shape[0]), _p(xb, c_float_p), C.c_long(nq), _p(xq, c_float_p), C.c_long(k),
                          _p(D, c_float_p), _p(I, c_long_p))
    return D, I


def ref_kmeans(x, k, niter=10, seed=1234):
    x = _c(x, _f32)
    n, d = x.shape
    cent = np.empty((k, d), _f32)
    ref().ref_kmeans(d, k, C.c_long(n), _p(x, c_float_p), niter, C.c_long(seed), _p(cent, c_float_p))
    return cent


def ref_rand_perm(n, seed):
    perm = np.empty(n, _i32)
    ref().ref_rand_perm(_p(perm, c_int_p), C.c_long(n), C.c_long(seed))
    return perm


def ref_pq_train(x, M, nbits=8):
    x = _c(x, _f32)
    n, d = x.shape
    cent = np.empty((M, 1 << nbits, d // M), _f32)
    ref().ref_pq_train(d, M, nbits, C.c_long(n), _p(x, c_float_p), _p(cent, c_float_p))
    return cent


def ref_pq_compute_codes(x, pq, nbits=8):
    x = _c(x, _f32)
    pq = _c(pq, _f32)
    n, d = x.shape
    M = pq.shape[0]
    codes = np.empty((n, M), _u8)
    ref().ref_pq_compute_codes(d, M, nbits, _p(pq, c_float_p), C.c_long(n), _p(x, c_float_p), _p(codes, c_u8_p))
    return codes


def ref_imi_search(x, cent, k):
    """the reference's MultiIndexQuantizer::search on the same sub-space codebooks"""
    x = _c(x, _f32)
    cent = _c(cent, _f32)
    n, d = x.shape
    M, ksub, _ = cent.shape
    nbits = int(round(np.log2(ksub)))
    assert 1 << nbits == ksub
    D = np.empty((n, k), _f32)
    I = np.empty((n, k), _i64)
    ref().ref_imi_search(d, M, nbits, _p(cent, c_float_p), C.c_long(n), _p(x, c_float_p), C.c_long(k),
                         _p(D, c_float_p), _p(I, c_long_p))
    return D, I


class RefIMIPQ:
    """the reference's IMI-PQ baseline index (tests/sift1b_imi_pq.cpp:216-236): MultiIndexQuantizer(d, 2, nbits_coarse)
    + IndexIVFPQ(2^(2 nbits_coarse) lists, M bytes per code), all reference CPU code"""

    def __init__(self, d, nbits_coarse, M, nbits=8):
        r = ref()
        r.ref_imipq_new.restype = C.c_void_p
        self.d, self.nbits_coarse, self.M, self.nbits = d, nbits_coarse, M, nbits
        self.h = C.c_void_p(r.ref_imipq_new(d, nbits_coarse, M, nbits))

    def codebooks(self):
        """-> coarse (2, K, d/2), pq (M, 2^nbits, d/M) of the trained index"""
        coarse = np.empty((2, 1 << self.nbits_coarse, self.d // 2), _f32)
        pq = np.empty((self.M, 1 << self.nbits, self.d // self.M), _f32)
        ref().ref_imipq_get_codebooks(self.h, _p(coarse, c_float_p), _p(pq, c_float_p))
        return coarse, pq

    def train(self, x):
        x = _c(x, _f32)
        ref().ref_imipq_train(self.h, C.c_long(x.shape[0]), _p(x, c_float_p))

    def add(self, x):
        x = _c(x, _f32)
        ref().ref_imipq_add(self.h, C.c_long(x.shape[0]), _p(x, c_float_p))

    def search(self, xq, k, nprobe):
        xq = _c(xq, _f32)
        D = np.empty((xq.shape[0], k), _f32)
        I = np.empty((xq.shape[0], k), _i64)
        ref().ref_imipq_search(self.h, C.c_long(xq.shape[0]), _p(xq, c_float_p), C.c_long(k), C.c_long(nprobe),
                               _p(D, c_float_p), _p(I, c_long_p))
        return D, I

    def __del__(self):
        try:
            if self.h:
                ref().ref_imipq_free(self.h)
                self.h = None
        except Exception:
            pass


def ref_heap_topk(vals, k):
    vals = _c(vals, _f32)
    n, m = vals.shape
    D = np.empty((n, k), _f32)
    I = np.empty((n, k), _i64)
    ref().ref_heap_topk(C.c_long(n), C.c_long(m), C.c_long(k), _p(vals, c_float_p), _p(D, c_float_p), _p(I, c_long_p))
    return D, I


class RefIVFPQ:
    """The reference's CPU IndexIVFPQ, optionally with externally supplied codebooks."""

    def __init__(self, d, nlist, M, nbits=8, coarse=None, pq=None, use_precomputed_table=1):
        self.d, self.nlist, self.M, self.nbits = d, nlist, M, nbits
        if coarse is not None:
            coarse = _c(coarse, _f32)
            pq = _c(pq, _f32)
        self.h = C.c_void_p(ref().ref_ivfpq_new(d, C.c_long(nlist), M, nbits, _p(coarse, c_float_p),
                                                _p(pq, c_float_p), use_precomputed_table))

    def train(self, x):
        x = _c(x, _f32)
        ref().ref_ivfpq_train(self.h, C.c_long(x.shape[0]), _p(x, c_float_p))

    def codebooks(self):
        coarse = np.empty((self.nlist, self.d), _f32)
        pq = np.empty((self.M, 1 << self.nbits, self.d // self.M), _f32)
        ref().ref_ivfpq_get_codebooks(self.h, _p(coarse, c_float_p), _p(pq, c_float_p))
        return coarse, pq

    def add(self, x):
        x = _c(x, _f32)
        ref().ref_ivfpq_add(self.h, C.c_long(x.shape[0]), _p(x, c_float_p))

    def search(self, xq, k, nprobe):
        xq = _c(xq, _f32)
        nq = xq.shape[0]
        D = np.empty((nq, k), _f32)
        I = np.empty((nq, k), _i64)
        ref().ref_ivfpq_search(self.h, C.c_long(nq), _p(xq, c_float_p), C.c_long(k), C.c_long(nprobe),
                               _p(D, c_float_p), _p(I, c_long_p))
        return D, I

    def get_list(self, l):
        n = ref().ref_ivfpq_list_size(self.h, C.c_long(l))
        ids = np.empty(n, _i64)
        codes = np.empty((n, self.M), _u8)
        if n:
            ref().ref_ivfpq_get_list(self.h, C.c_long(l), _p(ids, c_long_p), _p(codes, c_u8_p))
        return ids, codes

    def __del__(self):
        try:
            ref().ref_ivfpq_free(self.h)
        except Exception:
            pass


def ref_shards_flat_search(xb, shard_sizes, xq, k):
    xb = _c(xb, _f32)
    xq = _c(xq, _f32)
    sizes = _c(shard_sizes, _i64)
    nq, d = xq.shape
    D = np.empty((nq, k), _f32)
    I = np.empty((nq, k), _i64)
    ref().ref_shards_flat_search(d, len(sizes), _p(sizes, c_long_p), _p(xb, c_float_p), C.c_long(nq),
                                 _p(xq, c_float_p), C.c_long(k), _p(D, c_float_p), _p(I, c_long_p))
    return D, I
