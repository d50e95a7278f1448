"""Query path (a11-a15) against the oracle with PER-MISMATCH near-tie verification in float64 (tests/parity_util.py):
no fraction-based bars.  Geometries: SIFT-shaped d=128 / M=16 / E=32 (C2), DEEP-shaped d=96 / M=8 (C3), the 64-edge,
256-level graphs of the 1B drivers (gpu/test/sift1b16_query.cpp:252-254), and indexes with >= 100-entry lists on which
every scan kernel (flattened stream, warp-per-list, warp-autonomous, bank-skewed, long-list) is compared with the oracle
directly.  Both coarse routes: exact fp32 CUDA-core kernels and the tcgen05 route (fused top-P + line selection).
"""
import numpy as np
import pytest

from tests.parity_util import REL_TIE, Model64, check_lines, check_topk

pytestmark = pytest.mark.gpu


def T(a, dev):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops(cuda):
    from vector_line_quantization_b200 import ops as _ops

    return _ops


def _make_model(oracle, shape, d, C, E, M, nL, nt, nb, nq, kc, seed):
    from vector_line_quantization_b200 import data

    gen = data.sift_like if shape == "sift" else data.deep_like
    xt = gen(nt, d=d, kc=kc, seed=seed)
    m = oracle.train_all(xt, nlist=C, E=E, M=M, nL=nL, niter=5, pq_niter=5)
    m.update(xb=gen(nb, d=d, kc=kc, seed=seed + 1), xq=gen(nq, d=d, kc=kc, seed=seed + 2), d=d, C=C, E=E, M=M)
    return m


GEOMS = {
    # name: (shape, d, C, E, M, nL, n_train, n_base, nq, kc)
    "deep_d96_m8": ("deep", 96, 128, 32, 8, 256, 8192, 12000, 48, 256),
    "e64_l256": ("sift", 128, 128, 64, 16, 256, 8192, 12000, 48, 256),
    "long_lists_m16": ("sift", 128, 64, 8, 16, 64, 6000, 60000, 40, 128),    # 512 lists, ~117 entries each
    "long_lists_deep_m8": ("deep", 96, 64, 8, 8, 256, 6000, 60000, 40, 128),
}


@pytest.fixture(scope="module")
def models(oracle, small_model):
    out = {"c2_small": small_model}
    for i, (name, g) in enumerate(GEOMS.items()):
        out[name] = _make_model(oracle, *g, seed=100 + 10 * i)
    return out


_index_cache = {}


def _index(ops, cuda, oracle, name, m):
    """encode + list build on the device; the device-encoded entries are handed to the oracle, so both sides search the
    SAME index and only the query path differs"""
    import torch

    if name in _index_cache:
        return _index_cache[name]
    cent = T(m["cent"], cuda)
    cn = ops.row_norms(cent)
    edge, ed2, lcb, pq = T(m["edge"], cuda), T(m["edge_d2"], cuda), T(m["lambda_cb"], cuda), T(m["pq"], cuda)
    x = T(m["xb"], cuda)
    A, _ = ops.l2_assign(x, cent, cn)
    enc = ops.line_encode(x, A, cent, edge, ed2, lcb, pq)
    nl = m["C"] * m["E"]
    lists = ops.build_lists(nl, m["M"], enc.list, enc.codes, enc.lamq, enc.kappa,
                            torch.arange(x.shape[0], dtype=torch.int64, device=cuda))
    e_list, e_lamq, e_codes = N(enc.list), N(enc.lamq), N(enc.codes)
    off, perm = oracle.build_lists(e_list, nl)
    assert np.array_equal(N(lists.offsets), off) and np.array_equal(N(lists.ids), perm)
    gi = dict(cent=cent, cn=cn, edge=edge, ed2=ed2, lcb=lcb, pq=pq, lists=lists, e_list=e_list, e_lamq=e_lamq,
              e_codes=e_codes, off=off, perm=perm, m64=Model64(m))
    _index_cache[name] = gi
    return gi


def _oracle_search(oracle, m, gi, P, W, k, cap=1024):
    perm = gi["perm"]
    return oracle.search(m["xq"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"], gi["off"],
                         gi["e_codes"][perm], gi["e_lamq"][perm], perm.astype(np.int64), P=P, W=W, k=k, cap=cap,
                         want_debug=True)


def _check_coarse(cid, coarse_ref, xq, m64):
    cid, coarse_ref = np.asarray(cid), np.asarray(coarse_ref)
    qi, r = np.nonzero(cid != coarse_ref)
    if len(qi):
        q = np.asarray(xq, np.float64)[qi]
        dg, do = m64.coarse(q, cid[qi, r]), m64.coarse(q, coarse_ref[qi, r])
        assert np.all(np.abs(dg - do) <= REL_TIE * (np.abs(do) + (q ** 2).sum(1))), "unexplained top-P mismatch"
    return len(qi)


def _coarse_and_lines(ops, cuda, m, gi, P, W, route):
    import torch

    q = T(m["xq"], cuda)
    if route == "tc":
        pack = ops.CentPack(gi["cent"], gi["cn"])
        bm = torch.empty((q.shape[0], ops.num_buckets(m["C"])), dtype=torch.float32, device=cuda)
        D = ops.l2_distances_tc(q, pack, bucket_min=bm)
        lst, t1, t6, cid = ops.coarse_select_lines(D, bm, m["C"], min(P, m["C"]), gi["edge"], gi["ed2"], W, want_coarse=True)
    elif route == "tc_exact":  # no distance matrix: bucket minima + exact re-evaluation of the needed columns
        pack = ops.CentPack(gi["cent"], gi["cn"])
        if not ops._abi.lib().vlq_coarse_exact_supported(q.shape[1], m["C"], min(P, m["C"]), gi["edge"].shape[1], W):
            pytest.skip("shape outside the matrix-free kernel")
        bm = torch.empty((q.shape[0], ops.num_buckets(m["C"])), dtype=torch.float32, device=cuda)
        ops.l2_bucket_min_tc(q, pack, bm)
        lst, t1, t6, cid = ops.coarse_select_lines_exact(q, gi["cent"], gi["cn"], bm, min(P, m["C"]), gi["edge"], gi["ed2"],
                                                         W, want_coarse=True)
    else:
        D = ops.l2_distances(q, gi["cent"], gi["cn"])
        _, cid = ops.select_rows(D, min(P, m["C"]))
        lst, t1, t6 = ops.select_lines(D, cid, gi["edge"], gi["ed2"], W)
    return q, cid, lst, t1, t6


CASES = [
    # (model, P, W, k, cap)
    ("c2_small", 16, 128, 10, 1024), ("c2_small", 32, 256, 100, 1024), ("c2_small", 8, 1024, 1024, 1024),
    ("c2_small", 16, 64, 50, 2), ("c2_small", 1, 1, 1, 1024),
    ("deep_d96_m8", 16, 64, 100, 1024), ("deep_d96_m8", 32, 256, 10, 1024),
    ("e64_l256", 16, 256, 100, 1024), ("e64_l256", 8, 64, 20, 1024),
    ("long_lists_m16", 16, 64, 100, 1024), ("long_lists_m16", 32, 128, 128, 1024), ("long_lists_m16", 8, 32, 10, 50),
    ("long_lists_deep_m8", 16, 64, 100, 1024), ("long_lists_deep_m8", 64, 256, 200, 1024),
]


@pytest.mark.parametrize("route", ["simt", "tc", "tc_exact"])
@pytest.mark.parametrize("name,P,W,k,cap", CASES)
def test_query_path_vs_oracle_near_tie_verified(ops, cuda, oracle, models, name, P, W, k, cap, route):
    m = models[name]
    gi = _index(ops, cuda, oracle, name, m)
    Do, Io, coarse_ref, lines_ref, _ = _oracle_search(oracle, m, gi, P, W, k, cap)
    q, cid, lst, t1, t6 = _coarse_and_lines(ops, cuda, m, gi, P, W, route)
    _check_coarse(N(cid), coarse_ref, m["xq"], gi["m64"])
    # (1) the line lists: every rank that differs from the oracle's is a float64 line-score near-tie (the two directions of
    #     one graph edge score identically in exact arithmetic, so such ties are common at the W boundary)
    same_lines = check_lines(N(lst), lines_ref, m["xq"], gi["m64"])
    # (2) the scan, for ALL queries, against the oracle's scan of the SAME line choice (no query is exempt); where the
    #     line sets agree that is the oracle's own full search
    perm = gi["perm"]
    Dl, Il = oracle.scan_lines(m["xq"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"], gi["off"],
                               gi["e_codes"][perm], gi["e_lamq"][perm], perm.astype(np.int64), N(lst), k, cap)
    same_order = (N(lst) == lines_ref).all(axis=1)
    assert np.array_equal(Dl[same_order], Do[same_order]) and np.array_equal(Il[same_order], Io[same_order])
    ed2f = gi["ed2"].reshape(-1)
    avg_len = len(gi["perm"]) / (m["C"] * m["E"])
    # every scan kernel the dispatcher can choose, each against the oracle directly:
    #   hint 0 flattened stream; 30 warp-autonomous (k <= 128) or warp-per-list; 100 bank-skewed (M = 8, 16);
    #   400 the long-list kernel of scan_long.cu (M = 8, 16)
    for hint in (0, 30, 100, 400, int(avg_len)):
        D, I = ops.scan_topk(q, gi["pq"], gi["lcb"], lst, t1, t6, ed2f, gi["lists"], k, cap, list_len_hint=hint)
        check_topk(N(D), N(I), Dl, Il, m["xq"], gi["m64"], gi["e_list"], gi["e_lamq"], gi["e_codes"])
        check_topk(N(D), N(I), Do, Io, m["xq"], gi["m64"], gi["e_list"], gi["e_lamq"], gi["e_codes"], same_lines)
    D, I = ops.scan_topk(q, gi["pq"], gi["lcb"], lst, t1, t6, ed2f, gi["lists"], k, cap, use_workspace=False)
    check_topk(N(D), N(I), Dl, Il, m["xq"], gi["m64"], gi["e_list"], gi["e_lamq"], gi["e_codes"])


@pytest.mark.parametrize("name", ["c2_small", "deep_d96_m8", "e64_l256", "long_lists_m16"])
def test_search_entry_point_vs_oracle(ops, cuda, oracle, models, name):
    """ops.search (the tiled path bench.py times), both coarse routes, tile smaller than the batch"""
    m = models[name]
    gi = _index(ops, cuda, oracle, name, m)
    P, W, k = 16, 128, 50
    Do, Io, _, lines_ref, _ = _oracle_search(oracle, m, gi, P, W, k)
    for route in ("simt", "tc"):
        _, _, lst, _, _ = _coarse_and_lines(ops, cuda, m, gi, P, W, route)
        same_lines = check_lines(N(lst), lines_ref, m["xq"], gi["m64"])
        pack = ops.CentPack(gi["cent"], gi["cn"]) if route == "tc" else None
        D, I = ops.search(T(m["xq"], cuda), gi["cent"], gi["cn"], gi["edge"], gi["ed2"], gi["lcb"], gi["pq"], gi["lists"],
                          P, W, k, tile=17, pack=pack)
        check_topk(N(D), N(I), Do, Io, m["xq"], gi["m64"], gi["e_list"], gi["e_lamq"], gi["e_codes"], same_lines)
