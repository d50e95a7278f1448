"""N>1 host-side logic on CPU: world_size-2 gloo run of the shard / all-gather / merge protocol.
The per-shard search and the merge are done by the CPU oracle here (checker); on the GPU the same protocol runs with
NCCL + vlq_merge_topk (tests/test_gpu_parity.py::test_sharded_search_equals_single, bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, xb, xq, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from vector_line_quantization_b200 import sharding

    b, e = sharding.shard_range(len(xb), world, rank)

    def local_search(q, kk):
        D, I = po.l2_topk(q.numpy(), xb[b:e], kk, add_xnorm=True)
        return torch.from_numpy(D), torch.from_numpy(I.astype(np.int64) + b)  # global ids

    def merge(gD, gI):
        assert gD.shape == (world, len(xq), k)
        D, I = po.merge_topk(gD.numpy(), gI.numpy())
        return torch.from_numpy(D), torch.from_numpy(I)

    D, I = sharding.sharded_search(local_search, merge, torch.from_numpy(xq), k)
    ret[rank] = (D.numpy(), I.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from vector_line_quantization_b200 import sharding

    for n, w in [(10, 3), (1_000_000_007, 8), (5, 8)]:
        prev = 0
        for r in range(w):
            b, e = sharding.shard_range(n, w, r)
            assert b == prev and e >= b
            prev = e
        assert prev == n


def test_two_rank_sharded_search_equals_single():
    from oracle import pyoracle as po
    from vector_line_quantization_b200 import data

    po.lib()
    xb = data.deep_like(3001, seed=1)
    xq = data.deep_like(17, seed=2)
    k = 10
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), xb, xq, k, ret), nprocs=world, join=True)
    D0, I0 = po.l2_topk(xq, xb, k, add_xnorm=True)
    for r in range(world):  # every rank ends with the same merged result == the unsharded search
        D, I = ret[r]
        assert np.array_equal(D, D0)
        assert np.array_equal(I, I0.astype(np.int64))


def _worker_query_split(rank, world, port, m, enc, P, W, k, shard_by, ret):
    """The strong-scaling protocol of sharding.QuerySplitSearch with gloo collectives and the oracle as compute:
    rank r selects the lines of ITS query slice, the (list) slices are all-gathered, every rank scans ALL queries on its
    shard, the per-shard top-k are gathered and merged by query slice.  Shards: id ranges, or list ranges after
    sharding.route_by_list (variable-size all-to-all)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from vector_line_quantization_b200 import sharding

    lst, lamq, codes = enc
    n = len(lst)
    L = m["C"] * m["E"]
    b, e = sharding.shard_range(n, world, rank)
    my_list = torch.from_numpy(lst[b:e].copy())
    arrays = [torch.from_numpy(codes[b:e].copy()), torch.from_numpy(lamq[b:e].copy()),
              torch.arange(b, e, dtype=torch.int64)]
    if shard_by == "lists":
        my_list, arrays = sharding.route_by_list(my_list, arrays, L)
        own = (my_list.to(torch.int64) * world) // L
        assert bool((own == rank).all())  # every entry landed on the owner of its list
    off, perm = po.build_lists(my_list.numpy(), L)
    s_codes, s_lamq, s_ids = (a.numpy()[perm] for a in arrays)
    nq = len(m["xq"])
    q0 = [r * nq // world for r in range(world + 1)]
    args = (m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"])
    # 1. coarse stage + line selection for this rank's query slice only (it does not touch the lists)
    empty = (np.zeros(L + 1, np.int64), np.zeros((0, m["M"]), np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.int64))
    _, _, _, lines_mine, _ = po.search(m["xq"][q0[rank]:q0[rank + 1]], *args, *empty, P=P, W=W, k=k, want_debug=True)
    # 2. exchange the line slices (12 W bytes per query on the GPU: list, term1, term6; the oracle recomputes the terms)
    gathered = [torch.empty((q0[r + 1] - q0[r], W), dtype=torch.int32) for r in range(world)]
    dist.all_gather(gathered, torch.from_numpy(lines_mine.astype(np.int32)))
    lines = torch.cat(gathered).numpy()
    # 3. every rank scans ALL queries on its shard
    D, I = po.scan_lines(m["xq"], *args, off, s_codes, s_lamq, s_ids, lines, k, cap=1 << 20)
    gD, gI = sharding.gather_topk(torch.from_numpy(D), torch.from_numpy(I))
    # 4. merge by query slice
    s, t = q0[rank], q0[rank + 1]
    Dm, Im = po.merge_topk(gD[:, s:t].contiguous().numpy(), gI[:, s:t].contiguous().numpy())
    ret[rank] = (Dm, Im, int(len(my_list)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_query_split_search_equals_single(small_model, oracle):
    import pytest

    m = small_model
    A = oracle.l2_topk(m["xb"], m["cent"], 1)[1][:, 0].astype(np.int32)
    lst, lam = oracle.line_stage(m["xb"], A, m["cent"], m["edge"], m["edge_d2"])
    lamq = oracle.lambda_quantize(lam, m["lambda_cb"])
    r = oracle.residual(m["xb"], lst, lamq, m["lambda_cb"], m["cent"], m["edge"])
    codes = oracle.pq_encode(r, m["pq"])
    P, W, k, world = 16, 128, 10, 2
    off, perm = oracle.build_lists(lst, m["C"] * m["E"])
    D0, I0 = oracle.search(m["xq"], m["cent"], m["edge"], m["edge_d2"], m["lambda_cb"], m["pq"], off, codes[perm], lamq[perm],
                           perm.astype(np.int64), P=P, W=W, k=k, cap=1 << 20)
    mm = {key: m[key] for key in ("cent", "edge", "edge_d2", "lambda_cb", "pq", "xq", "C", "E", "M")}
    for shard_by in ("ids", "lists"):
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker_query_split, args=(world, _free_port(), mm, (lst, lamq, codes), P, W, k, shard_by, ret),
                 nprocs=world, join=True)
        nq = len(m["xq"])
        D = np.concatenate([ret[r_][0] for r_ in range(world)])
        I = np.concatenate([ret[r_][1] for r_ in range(world)])
        assert sum(ret[r_][2] for r_ in range(world)) == len(lst)
        assert D.shape == (nq, k)
        assert np.array_equal(D, D0)  # same entries, same arithmetic: bit-identical distances
        assert (I == I0).mean() > 0.999  # ids may swap on exact distance ties only
