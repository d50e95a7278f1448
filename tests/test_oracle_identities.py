"""The VLQ-only arithmetic has no CPU twin in the reference; these identities pin the oracle's restatement of it.  CPU only."""
import numpy as np

from vector_line_quantization_b200 import data


def test_line_stage_is_nearest_line(oracle, small_model):
    m = small_model
    x = m["xb"][:3000]
    _, A = oracle.l2_topk(x, m["cent"], 1)
    A = A[:, 0].copy()
    lst, lam = oracle.line_stage(x, A, m["cent"], m["edge"], m["edge_d2"])
    E = m["E"]
    assert np.array_equal(lst // E, A)
    e = lst % E
    x64, c64 = x.astype(np.float64), m["cent"].astype(np.float64)
    for i in range(0, 3000, 37):
        a = c64[A[i]]
        best_q, best_e, best_valid = None, None, False
        for ee in range(E):
            s = c64[m["edge"][A[i], ee]]
            t = np.dot(x64[i] - a, s - a) / np.dot(s - a, s - a)  # true projection parameter
            q2 = np.sum((x64[i] - (a + t * (s - a))) ** 2)
            valid = 0.0 <= t <= 1.0
            better = best_q is None or (valid and not best_valid) or (valid == best_valid and q2 < best_q)
            if better and not (best_valid and not valid):
                best_q, best_e, best_valid = q2, ee, valid
        s = c64[m["edge"][A[i], e[i]]]
        t = np.dot(x64[i] - a, s - a) / np.dot(s - a, s - a)
        q2 = np.sum((x64[i] - (a + t * (s - a))) ** 2)
        assert abs(t - lam[i]) < 2e-3 * max(1.0, abs(t))
        assert q2 <= best_q * (1 + 1e-4) + 1e-6


def test_lambda_quantizer_is_nearest_level(oracle):
    rng = np.random.RandomState(3)
    cb = rng.rand(256).astype(np.float32)
    lam = rng.normal(0.3, 0.5, 5000).astype(np.float32)
    q = oracle.lambda_quantize(lam, cb)
    want = np.argmin((lam[:, None] - cb[None, :]) ** 2, axis=1)
    assert np.array_equal(q, want)


def test_scan_distance_is_decode_distance(oracle, small_model):
    """dist == ||q - ((1-l)c + l s) - p(code)||^2 - ||q||^2"""
    m = small_model
    enc = oracle.encode_all(m["xb"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"])
    nl = m["C"] * m["E"]
    offsets, perm = oracle.build_lists(enc["list"], nl)
    k = 20
    D, I = oracle.search(m["xq"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"], offsets,
                         enc["codes"][perm], enc["lamq"][perm], perm.astype(np.int64), P=16, W=128, k=k)
    assert (I >= 0).all()
    for qi in range(0, m["xq"].shape[0], 5):
        for r in range(0, k, 3):
            ent = I[qi, r]
            lst = enc["list"][ent]
            c, e = lst // m["E"], lst % m["E"]
            s = m["edge"][c, e]
            want = oracle.decode_distance(m["xq"][qi], m["cent"][c], m["cent"][s], m["lambda_cb"][enc["lamq"][ent]],
                                          m["pq"], enc["codes"][ent])
            assert abs(D[qi, r] - want) <= 1e-4 * abs(want) + 1e-2
        assert np.all(np.diff(D[qi]) >= 0)


def test_search_recall_sane(oracle, small_model):
    m = small_model
    enc = oracle.encode_all(m["xb"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"])
    offsets, perm = oracle.build_lists(enc["list"], m["C"] * m["E"])
    _, I = oracle.search(m["xq"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"], offsets,
                         enc["codes"][perm], enc["lamq"][perm], perm.astype(np.int64), P=32, W=256, k=100)
    _, gt = oracle.l2_topk(m["xq"], m["xb"], 1)
    assert data.recall_at(I, gt[:, 0], 100) > 0.8


def test_merge_equals_global_topk(oracle):
    rng = np.random.RandomState(9)
    R, nq, k = 4, 30, 16
    D = np.sort(rng.rand(R, nq, k).astype(np.float32), axis=2)
    I = rng.randint(0, 1 << 40, size=(R, nq, k)).astype(np.int64)
    oD, oI = oracle.merge_topk(D, I)
    flatD = D.transpose(1, 0, 2).reshape(nq, R * k)
    flatI = I.transpose(1, 0, 2).reshape(nq, R * k)
    order = np.argsort(flatD, axis=1, kind="stable")[:, :k]
    assert np.array_equal(oD, np.take_along_axis(flatD, order, 1))
    assert np.array_equal(oI, np.take_along_axis(flatI, order, 1))
