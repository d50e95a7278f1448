"""Pins oracle/vlq_oracle.c against the UNMODIFIED reference CPU library (oracle/_ref/libfaiss_ref.so, built from the
sources under /root/reference by oracle/Makefile).  CPU only.  Skipped when the prebuilt reference library is absent.

What the reference can pin: RNG/permutation, coarse assignment, k-means, PQ encode, heap top-k, shard
merge, and the lambda == 0 slice of the VLQ search (== IndexIVFPQ precomputed-table search).  The VLQ-only arithmetic
has no CPU implementation in the reference; it is pinned by identities in test_oracle_identities.py.
"""
import numpy as np
import pytest

from vector_line_quantization_b200 import data

REL_TIE = 1e-5  # north_star: ids identical except near-ties with relative distance gap < 1e-5


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libfaiss_ref.so not built")
    oracle.ref()
    return oracle


def ids_equal_up_to_ties(ids_a, ids_b, x, cent, rel=REL_TIE):
    """all mismatching rows must be near-ties: |d(x,c_a) - d(x,c_b)| <= rel * d"""
    bad = np.nonzero(ids_a != ids_b)[0]
    for i in bad:
        da = np.sum((x[i].astype(np.float64) - cent[ids_a[i]]) ** 2)
        db = np.sum((x[i].astype(np.float64) - cent[ids_b[i]]) ** 2)
        assert abs(da - db) <= rel * max(da, db) + 1e-12, (i, da, db)
    return len(bad)


def test_rand_perm_matches_reference(ref):
    for n, seed in [(1, 5), (17, 1234), (1000, 1235), (65536, 7)]:
        assert np.array_equal(ref.rand_perm(n, seed), ref.ref_rand_perm(n, seed))


def test_coarse_assignment_matches_indexflat(ref):
    x = data.sift_like(3000, seed=11)
    cent = data.sift_like(700, seed=12) + np.float32(0.25)
    D, I = ref.l2_topk(x, cent, 1, add_xnorm=True)
    Dr, Ir = ref.ref_flat_search(cent, x, 1)
    nbad = ids_equal_up_to_ties(I[:, 0], Ir[:, 0].astype(np.int32), x, cent)
    assert nbad <= 3
    ok = I[:, 0] == Ir[:, 0]
    np.testing.assert_allclose(D[ok, 0], Dr[ok, 0], rtol=1e-4)


def test_topk_matches_indexflat(ref):
    x = data.deep_like(300, seed=21)
    base = data.deep_like(5000, seed=22)
    for k in (1, 10, 100):
        D, I = ref.l2_topk(x, base, k, add_xnorm=True)
        Dr, Ir = ref.ref_flat_search(base, x, k)
        np.testing.assert_allclose(D, Dr, rtol=2e-4, atol=2e-6)
        assert (I == Ir).mean() > 0.99


def test_heap_topk_semantics(ref):
    rng = np.random.RandomState(0)
    vals = rng.rand(50, 3000).astype(np.float32)
    vals[:, ::7] = vals[:, 1::7]  # duplicates: ties
    for k in (1, 5, 128, 1024):
        Dr, _ = ref.ref_heap_topk(vals, k)
        want = np.sort(vals, axis=1)[:, :k]
        assert np.array_equal(Dr, want)


def test_kmeans_matches_clustering(ref):
    x = data.sift_like(6000, kc=64, seed=31)
    cr = ref.ref_kmeans(x, 40, niter=5, seed=1234)
    co, _ = ref.kmeans(x, 40, niter=5, seed=1234)
    # identical init (same RNG) and same update rule; assignment near-ties may move single points
    np.testing.assert_allclose(co, cr, rtol=2e-3, atol=0.5)
    assert np.mean(np.abs(co - cr) < 1e-2) > 0.95


def test_kmeans_subsample_path(ref):
    x = data.sift_like(3000, kc=16, seed=33)
    k = 8  # n > k*256 -> the reference subsamples with rand_perm(seed)
    cr = ref.ref_kmeans(x, k, niter=4, seed=99)
    co, _ = ref.kmeans(x, k, niter=4, seed=99)
    np.testing.assert_allclose(co, cr, rtol=2e-3, atol=0.5)


def test_pq_encode_matches_product_quantizer(ref):
    rng = np.random.RandomState(5)
    for d, M in [(128, 16), (96, 8), (64, 4)]:
        r = rng.normal(0, 20, size=(2000, d)).astype(np.float32)
        pq = ref.ref_pq_train(r[:1500], M)
        codes_ref = ref.ref_pq_compute_codes(r, pq)
        codes = ref.pq_encode(r, pq)
        mism = np.argwhere(codes != codes_ref)
        dsub = d // M
        for i, m in mism:  # only exact/near ties may differ (BLAS path for dsub >= 16)
            sub = r[i, m * dsub:(m + 1) * dsub].astype(np.float64)
            da = np.sum((sub - pq[m, codes[i, m]]) ** 2)
            db = np.sum((sub - pq[m, codes_ref[i, m]]) ** 2)
            assert abs(da - db) <= 1e-5 * max(da, db)
        assert len(mism) <= 5


def test_shard_merge_matches_indexshards(ref):
    xb = data.deep_like(4000, seed=41)
    xq = data.deep_like(50, seed=42)
    sizes = [1000, 1500, 700, 800]
    k = 20
    Dr, Ir = ref.ref_shards_flat_search(xb, sizes, xq, k)
    Ds, Is = [], []
    off = 0
    for s in sizes:
        D, I = ref.ref_flat_search(xb[off:off + s], xq, k)
        Ds.append(D)
        Is.append(I + off)
        off += s
    Dm, Im = ref.merge_topk(np.stack(Ds), np.stack(Is))
    assert np.array_equal(Dm, Dr)
    assert (Im == Ir).mean() > 0.999  # ties between shards may order differently


def test_lambda0_search_equals_indexivfpq(ref):
    """With a single lambda level == 0 the VLQ anchor is the centroid itself, so searching all lines of the P probed
    centroids must reproduce IndexIVFPQ (precomputed tables) with nprobe = P."""
    d, C, E, M, P, k = 64, 32, 8, 8, 6, 10
    xt = data.sift_like(5000, d=d, kc=64, seed=51)
    xb = data.sift_like(6000, d=d, kc=64, seed=52)
    xq = data.sift_like(40, d=d, kc=64, seed=53)
    ivf = ref.RefIVFPQ(d, C, M)
    ivf.train(xt)
    cent, pq = ivf.codebooks()
    ivf.add(xb)
    Dr, Ir = ivf.search(xq, k, P)

    edge, ed2 = ref.knn_graph(cent, E)
    lcb = np.zeros(1, np.float32)
    enc = ref.encode_all(xb, cent, edge, ed2, lcb, pq)
    # PQ codes of the residual x - c_A must equal what IndexIVFPQ stored
    for l in range(C):
        ids_l, codes_l = ivf.get_list(l)
        assert np.array_equal(enc["codes"][ids_l], codes_l)
        assert np.all(enc["A"][ids_l] == l)
    offsets, perm = ref.build_lists(enc["list"], C * E)
    D, I = ref.search(xq, cent, edge, ed2, lcb, pq, offsets, enc["codes"][perm], enc["lamq"][perm], perm.astype(np.int64),
                      P=P, W=P * E, k=k, cap=1 << 20)
    qn = np.sum(xq.astype(np.float64) ** 2, axis=1, keepdims=True)
    np.testing.assert_allclose(D + qn, Dr, rtol=2e-4)
    assert (I == Ir).mean() > 0.97


@pytest.mark.parametrize("d,M,nbits,k", [(32, 2, 6, 1), (32, 2, 6, 50), (64, 2, 8, 300), (48, 3, 4, 100), (128, 2, 8, 2048)])
def test_imi_search_matches_multi_index_quantizer(ref, d, M, nbits, k):
    """next row f3: the oracle's IMI coarse quantizer (multi-sequence walk) against MultiIndexQuantizer::search
    (IndexPQ.cpp:813-855) on the same sub-space codebooks.

    Reference quirk (documented deviation): MinSumK runs with use_seen = false and recognises a cell that was pushed
    along two paths only when both copies surface at the top of the heap together (IndexPQ.cpp:724-733); the two
    copies carry sums accumulated in different orders, so they can differ in the last bit, and then the SAME cell is
    returned twice and the list holds fewer than k distinct cells.  The oracle returns the k distinct cells with the
    smallest sums (checked against brute force below); the reference's list, de-duplicated, must be its prefix."""
    rng = np.random.RandomState(d + k)
    ksub = 1 << nbits
    cent = rng.rand(M, ksub, d // M).astype(np.float32)
    x = rng.rand(40, d).astype(np.float32)
    k = min(k, ksub ** M)
    D, I = ref.imi_search(x, cent, k)
    Dr, Ir = ref.ref_imi_search(x, cent, k)
    assert np.all(np.diff(D, axis=1) >= 0)
    dsub = d // M
    tabs = [((x[:, None, m * dsub:(m + 1) * dsub] - cent[m][None]) ** 2).sum(-1) for m in range(M)]  # [n][ksub] each
    for i in range(x.shape[0]):
        assert len(set(I[i])) == k  # no cell twice
        # brute force over the whole grid: the k smallest sums
        grid = tabs[0][i]
        for m in range(1, M):
            grid = (grid[None, :] + tabs[m][i][:, None]).reshape(-1)  # label = sum_m j_m * ksub^m
        order = np.argsort(grid, kind="stable")[:k]
        np.testing.assert_allclose(D[i], grid[order], rtol=2e-5, atol=1e-6)
        assert np.mean(I[i] == order) > 0.97  # equal sums may swap
        np.testing.assert_allclose(grid[I[i]], D[i], rtol=2e-5, atol=1e-6)  # every label decodes to its reported sum
        # the reference, first occurrences only, is a prefix of the oracle's list (up to swaps between near-equal sums)
        _, first = np.unique(Ir[i], return_index=True)
        keep = np.sort(first)
        ur, udr = Ir[i][keep], Dr[i][keep]
        n_u = len(ur)
        np.testing.assert_allclose(udr, D[i][:n_u], rtol=3e-5, atol=2e-6)
        assert np.mean(ur == I[i][:n_u]) > 0.97
        assert k - n_u <= max(2, k // 50)  # duplicates are rare


def test_reference_imipq_baseline_index(ref):
    """the IMI-PQ baseline of BASELINE configs[4] (tests/sift1b_imi_pq.cpp) builds and searches through the reference's
    own classes: exact-duplicate queries find themselves, results ascend, recall grows with nprobe"""
    rng = np.random.RandomState(4)
    d = 32
    xt = rng.rand(4000, d).astype(np.float32)
    xb = rng.rand(3000, d).astype(np.float32)
    xq = xb[:100] + rng.normal(0, 0.01, (100, d)).astype(np.float32)
    idx = ref.RefIMIPQ(d, 4, 8)  # 2 x 2^4 coarse codebooks = 256 cells, 8 bytes per code
    idx.train(xt)
    idx.add(xb)
    gt = ref.l2_topk(xq, xb, 1)[1][:, 0]
    rec = []
    for nprobe in (1, 8, 64):
        D, I = idx.search(xq, 10, nprobe)
        for i in range(100):
            assert np.all(np.diff(D[i][I[i] >= 0]) >= 0)
        rec.append(np.mean([gt[i] in I[i] for i in range(100)]))
    assert rec[0] <= rec[1] <= rec[2] and rec[2] > 0.5
