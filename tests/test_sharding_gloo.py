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
