"""tcgen05 coarse-assignment kernels (csrc/assign_tc.cu) against the exact fp32 CUDA-core path and the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from vector_line_quantization_b200 import ops as _ops

    return _ops


def _data(n, C, d, seed, kind):
    rng = np.random.RandomState(seed)
    if kind == "sift":
        x = np.clip(np.rint(rng.normal(60, 40, (n, d))), 0, 255).astype(np.float32)
        c = (np.clip(rng.normal(60, 40, (C, d)), 0, 255) + rng.rand(C, d)).astype(np.float32)
    else:  # deep-like: unit vectors
        x = rng.normal(size=(n, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        c = rng.normal(size=(C, d)).astype(np.float32)
        c /= np.linalg.norm(c, axis=1, keepdims=True) * 1.1
    return x, c


@pytest.mark.parametrize("n,C,d,kind", [(256, 128, 128, "sift"), (1000, 1000, 128, "sift"), (5000, 4096, 96, "deep"),
                                        (77, 333, 64, "deep"), (3000, 65536, 128, "sift"), (70000, 2048, 32, "sift")])
def test_assign_tc_matches_exact(ops, cuda, oracle, n, C, d, kind):
    import torch

    x, c = _data(n, C, d, n + C, kind)
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    pack = ops.CentPack(ct)
    ids, dist = ops.l2_assign_tc(xt, pack, add_xnorm=True)
    ids0, dist0 = ops.l2_assign(xt, ct, pack.cnorm, add_xnorm=True)
    torch.cuda.synchronize()
    ids, dist, ids0, dist0 = ids.cpu().numpy(), dist.cpu().numpy(), ids0.cpu().numpy(), dist0.cpu().numpy()
    bad = np.nonzero(ids != ids0)[0]
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    for i in bad:  # only near-ties may differ (north_star: relative distance gap < 1e-5)
        da, db = np.sum((x64[i] - c64[ids[i]]) ** 2), np.sum((x64[i] - c64[ids0[i]]) ** 2)
        assert abs(da - db) <= 1e-5 * max(da, db) + 1e-12, (i, da, db)
    assert len(bad) <= max(1, n // 2000)
    ok = ids == ids0
    scale = np.sum(x64 ** 2, axis=1)[ok] + np.sum(c64[ids0[ok]] ** 2, axis=1)
    assert np.all(np.abs(dist[ok] - dist0[ok]) <= 2e-6 * scale + 1e-6)
    if n * C <= 5_000_000:
        _, Io = oracle.l2_topk(x, c, 1)
        assert (ids == Io[:, 0]).mean() > 0.999


@pytest.mark.parametrize("n,C,d", [(300, 1000, 128), (1024, 4096, 96), (50, 130, 64)])
def test_distances_tc_matches_exact(ops, cuda, n, C, d):
    import torch

    x, c = _data(n, C, d, 7, "sift")
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    pack = ops.CentPack(ct)
    D = ops.l2_distances_tc(xt, pack)
    D0 = ops.l2_distances(xt, ct, pack.cnorm)
    ref = (c.astype(np.float64) ** 2).sum(1)[None, :] - 2 * x.astype(np.float64) @ c.astype(np.float64).T
    mag = np.linalg.norm(x, axis=1)[:, None] * np.linalg.norm(c, axis=1)[None, :] + 1
    err_tc = np.abs(D.cpu().numpy() - ref) / mag
    err_simt = np.abs(D0.cpu().numpy() - ref) / mag
    assert err_tc.max() < 3e-6, err_tc.max()       # fp32-GEMM grade
    assert err_tc.max() < 8 * err_simt.max() + 1e-7


def test_tc_rejects_unsupported_shapes(ops, cuda):
    import torch

    with pytest.raises(ValueError):
        ops.CentPack(torch.randn(100, 20, device=cuda))
    with pytest.raises(ValueError):
        ops.CentPack(torch.randn(100, 256, device=cuda))


@pytest.mark.parametrize("nq,C,d,P,W,E", [(100, 4096, 128, 64, 256, 32), (33, 1000, 96, 16, 100, 8), (10, 130, 64, 200, 1024, 16),
                                          (7, 65536, 128, 64, 1024, 32),
                                          # register-resident fast path with 16 keys per thread; W larger than the lines
                                          (20, 65536, 128, 128, 512, 32), (9, 8192, 128, 64, 700, 64), (5, 2048, 96, 4, 1024, 32),
                                          (6, 4096, 128, 1, 1, 32)])
def test_fused_coarse_select_equals_two_step(ops, cuda, nq, C, d, P, W, E):
    """bucket-minimum route == select_rows(k=P) + select_lines on the same distance matrix, bit for bit"""
    import torch

    x, c = _data(nq, C, d, 11, "sift")
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    pack = ops.CentPack(ct)
    bm = torch.empty((nq, ops.num_buckets(C)), dtype=torch.float32, device=cuda)
    D = ops.l2_distances_tc(xt, pack, bucket_min=bm)
    # bucket minima really are the minima of the 32-column buckets
    Dp = torch.full((nq, ops.num_buckets(C) * 32), float("inf"), device=cuda)
    Dp[:, :C] = D
    assert torch.equal(bm, Dp.view(nq, -1, 32).min(dim=2).values)
    g = torch.Generator(device="cpu").manual_seed(5)
    edge = torch.stack([torch.randperm(C, generator=g)[:E] for _ in range(C)]).to(torch.int32).to(cuda) if C <= 4096 else \
        torch.randint(0, C, (C, E), generator=g, dtype=torch.int32).to(cuda)
    ed2 = (torch.rand((C, E), generator=g) * 1000 + 1).to(cuda)
    Pk = min(P, C)
    _, cid = ops.select_rows(D, Pk)
    if Pk < P:
        cid = torch.cat([cid, torch.full((nq, P - Pk), -1, dtype=torch.int32, device=cuda)], dim=1).contiguous()
    l0, a0, b0 = ops.select_lines(D, cid, edge, ed2, W)
    l1, a1, b1, cid1 = ops.coarse_select_lines(D, bm, C, P, edge, ed2, W, want_coarse=True)
    assert torch.equal(cid1, cid)
    assert torch.equal(l1, l0) and torch.equal(a1, a0) and torch.equal(b1, b0)


@pytest.mark.parametrize("n,C,d,kind", [(600, 200, 128, "sift"), (5000, 4096, 96, "deep"), (260, 130, 64, "deep")])
def test_assign_tc_rows_of_any_magnitude(ops, cuda, n, C, d, kind):
    """rows far larger (x 1e4 .. 1e30) or smaller (x 1e-20) than the centroids: the per-row power-of-two pre-scale keeps
    every row inside fp16's range, so ids and distances stay fp32-grade (a fixed centroid-derived scale overflowed to
    inf for components > ~100 x the largest centroid component); exact ties resolve to the lowest id"""
    import torch

    x, c = _data(n, C, d, 3 * n + C, kind)
    c[C // 2] = c[3]  # an exact tie between two centroids
    x[5] = c[3] * 0.999
    x[6] = x[5]
    for i, f in ((10, 1e4), (11, 3e7), (12, 1e30), (13, 1e-20), (14, -2e5)):
        x[i] *= np.float32(f)
    x[15] = 0.0
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    pack = ops.CentPack(ct)
    ids, dist = ops.l2_assign_tc(xt, pack, add_xnorm=False)
    ids0, dist0 = ops.l2_assign(xt, ct, pack.cnorm, add_xnorm=False)
    Dm = ops.l2_distances_tc(xt, pack)
    D0 = ops.l2_distances(xt, ct, pack.cnorm)
    torch.cuda.synchronize()
    ids, ids0, dist, dist0 = ids.cpu().numpy(), ids0.cpu().numpy(), dist.cpu().numpy(), dist0.cpu().numpy()
    assert (ids >= 0).all() and np.isfinite(dist).all()
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    for i in np.nonzero(ids != ids0)[0]:
        da, db = np.sum((x64[i] - c64[ids[i]]) ** 2), np.sum((x64[i] - c64[ids0[i]]) ** 2)
        assert abs(da - db) <= 1e-5 * max(da, db) + 1e-12, (i, da, db)
    assert ids[5] == 3 and ids[6] == 3
    ref = (c64 ** 2).sum(1)[None, :] - 2 * x64 @ c64.T
    mag = np.linalg.norm(x64, axis=1)[:, None] * np.linalg.norm(c64, axis=1)[None, :] + (c64 ** 2).sum(1)[None, :]
    assert (np.abs(Dm.cpu().numpy() - ref) / mag).max() < 3e-6
    assert (np.abs(D0.cpu().numpy() - ref) / mag).max() < 3e-6


@pytest.mark.parametrize("nq,C,d,P,W,E", [(100, 4096, 128, 8, 256, 32), (64, 65536, 128, 8, 256, 32), (33, 1000, 96, 16, 100, 8),
                                          (10, 130, 64, 120, 1024, 16), (20, 65536, 128, 128, 512, 32),
                                          (9, 8192, 128, 64, 700, 64), (5, 2048, 96, 4, 1024, 32), (6, 4096, 128, 1, 1, 32)])
def test_matrix_free_coarse_select_against_float64(ops, cuda, nq, C, d, P, W, E):
    """vlq_l2_bucket_min_tc + vlq_coarse_select_lines_exact (no distance matrix): the top-P centroids and the W lines are
    the float64 ones up to near-ties (1e-5 of the magnitude of the terms), term1 / term6 are D[c] and D[s] - D[c]"""
    import torch

    x, c = _data(nq, C, d, 11, "sift")
    xt, ct = torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda)
    pack = ops.CentPack(ct)
    assert ops._abi.lib().vlq_coarse_exact_supported(d, C, min(P, C), E, W)
    bm = torch.empty((nq, ops.num_buckets(C)), dtype=torch.float32, device=cuda)
    ops.l2_bucket_min_tc(xt, pack, bm)
    bm1 = torch.empty_like(bm)
    ops.l2_distances_tc(xt, pack, bucket_min=bm1)
    assert torch.equal(bm, bm1)  # the matrix-free sweep writes the bucket minima of the matrix sweep
    g = torch.Generator(device="cpu").manual_seed(5)
    edge = torch.randint(0, C, (C, E), generator=g, dtype=torch.int32).to(cuda)
    ed2 = (torch.rand((C, E), generator=g) * 1000 + 1).to(cuda)
    Pk = min(P, C)
    lst, t1, t6, cid = ops.coarse_select_lines_exact(xt, ct, pack.cnorm, bm, Pk, edge, ed2, W, want_coarse=True)
    x64, c64 = xt.double(), ct.double()
    D = (c64 * c64).sum(1)[None, :] - 2.0 * x64 @ c64.T  # [nq][C] float64
    mag = (c64 * c64).sum(1)[None, :] + 2.0 * (x64.abs() @ c64.abs().T)
    tol = 1e-5 * mag.max(dim=1).values  # per query
    cid64 = cid.long()
    assert int((cid64 < 0).sum()) == 0 and all(len(set(r.tolist())) == Pk for r in cid64.cpu())
    dsel = torch.gather(D, 1, cid64)
    assert bool((dsel[:, 1:] - dsel[:, :-1] >= -tol[:, None]).all())  # ascending
    kth = torch.topk(D, Pk, dim=1, largest=False).values[:, -1]
    assert bool((dsel.max(dim=1).values <= kth + tol).all())  # the selected set is a top-P set up to near-ties
    # lines of the kernel's own top-P
    s = edge.long()[cid64]  # [nq][Pk][E]
    a2 = torch.gather(D, 1, s.reshape(nq, -1)).reshape(nq, Pk, E)
    b2 = dsel[:, :, None].expand(-1, -1, E)
    c2 = ed2.double()[cid64]
    v = a2 - b2 - c2
    score = torch.where(v > 0, b2, b2 - 0.25 * v * v / c2).reshape(nq, -1)
    lid = (cid64[:, :, None] * E + torch.arange(E, device=cuda)[None, None, :]).reshape(nq, -1)
    Wk = min(W, Pk * E)
    assert int((lst[:, :Wk] < 0).sum()) == 0 and int((lst[:, Wk:] >= 0).sum()) == 0
    pos = {}
    for r in range(nq):
        where = {int(l): i for i, l in enumerate(lid[r].tolist())}
        got = [where[int(l)] for l in lst[r, :Wk].tolist()]
        assert len(set(got)) == Wk
        sg = score[r, got]
        stol = 4 * tol[r]
        assert bool((sg[1:] - sg[:-1] >= -stol).all())
        assert float(sg.max()) <= float(torch.topk(score[r], Wk, largest=False).values[-1]) + float(stol)
        torch.testing.assert_close(t1[r, :Wk].double(), b2.reshape(nq, -1)[r, got], rtol=0, atol=float(tol[r]))
        torch.testing.assert_close(t6[r, :Wk].double(), (a2 - b2).reshape(nq, -1)[r, got], rtol=0, atol=float(2 * tol[r]))
    # NaN / inf queries select nothing
    xt2 = xt.clone()
    xt2[0, 3] = float("nan")
    ops.l2_bucket_min_tc(xt2, pack, bm)
    lst2, _, _, cid2 = ops.coarse_select_lines_exact(xt2, ct, pack.cnorm, bm, Pk, edge, ed2, W, want_coarse=True)
    assert int((lst2[0] >= 0).sum()) == 0 and int((cid2[0] >= 0).sum()) == 0
    assert torch.equal(lst2[1:], lst[1:])
