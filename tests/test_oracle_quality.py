"""The reference's own unit tests are statistical: tests/test_ivfpq_codec.cpp:27-65 asserts that
the reconstruction error falls with more centroids / more code bytes, tests/test_ivfpq_indexing.cpp:20-100 that a small
index finds the true neighbour for > 40 % of 200 queries.  The same two checks on the oracle's VLQ index (CPU only)."""
import numpy as np
import pytest


def _data(n, d, seed):
    rng = np.random.RandomState(seed)  # the reference tests use uniform drand48 vectors
    return rng.rand(n, d).astype(np.float32)


def _reconstruction_error(oracle, x, model):
    enc = oracle.encode_all(x, model["cent"], model["edge"], model["edge_d2"], model["lambda_cb"], model["pq"])
    M, _, dsub = model["pq"].shape
    p = np.concatenate([model["pq"][m][enc["codes"][:, m]] for m in range(M)], axis=1)
    return float(((enc["residual"] - p) ** 2).sum())


@pytest.mark.timeout(300)
def test_codec_error_falls_with_centroids_and_code_bytes(oracle):
    d, nt, nb = 64, 6000, 2000
    xt, xb = _data(nt, d, 1), _data(nb, d, 2)
    err = {}
    for nlist, M in ((16, 4), (64, 4), (16, 8)):
        model = oracle.train_all(xt, nlist, 8, M, 16, niter=6, pq_niter=8)
        err[(nlist, M)] = _reconstruction_error(oracle, xb, model)
    assert err[(16, 4)] > err[(64, 4)]  # more coarse centroids (and lines) -> smaller residuals
    assert err[(16, 4)] > err[(16, 8)]  # more PQ bytes -> smaller error


@pytest.mark.timeout(300)
def test_indexing_finds_true_neighbour(oracle):
    d, nt, nb, nq, E, M, k = 64, 4000, 3000, 200, 8, 16, 5
    rng = np.random.RandomState(35)
    xt, xb = rng.rand(nt, d).astype(np.float32), rng.rand(nb, d).astype(np.float32)
    xq = xb[:nq] + rng.normal(0, 0.01, (nq, d)).astype(np.float32)  # perturbed database vectors, as a sanity workload
    model = oracle.train_all(xt, 25, E, M, 16, niter=6, pq_niter=8)
    enc = oracle.encode_all(xb, model["cent"], model["edge"], model["edge_d2"], model["lambda_cb"], model["pq"])
    off, order = oracle.build_lists(enc["list"], 25 * E)
    D, I = oracle.search(xq, model["cent"], model["edge"], model["edge_d2"], model["lambda_cb"], model["pq"], off,
                         enc["codes"][order], enc["lamq"][order], order.astype(np.int64), P=5, W=20, k=k)
    gt = oracle.l2_topk(xq, xb, 1)[1][:, 0]
    found = np.mean([gt[i] in I[i] for i in range(nq)])
    assert found > 0.4  # tests/test_ivfpq_indexing.cpp:97
    assert np.all(np.diff(D, axis=1) >= 0)
