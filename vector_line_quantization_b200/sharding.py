"""Multi-GPU plumbing of the search path: id-range database shards, replicated queries, one exchange.

One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.  Every rank holds
rows [shard_range) of the database with GLOBAL ids, searches all queries on its shard, and the per-shard top-k lists
are all-gathered into the [rank][nq][k] layout the merge kernel (vlq_merge_topk == the reference's mergekernel,
gpu/GpuIndexIVFPQ.cu:1491-1515) consumes.
"""
import torch
import torch.distributed as dist


def shard_range(n_total, world, rank):
    """rows [b, e) of shard `rank`: contiguous id ranges, sizes differ by at most one (IndexShards::add,
    MetaIndexes.cpp:402-440 uses the same i*n/ns split)"""
    return rank * n_total // world, (rank + 1) * n_total // world


def gather_topk(D, I, out_D=None, out_I=None):
    """all-gather per-shard results (nq,k) -> ([world][nq][k], [world][nq][k]) on every rank"""
    world = dist.get_world_size()
    nq, k = D.shape
    if out_D is None:
        out_D = torch.empty((world, nq, k), dtype=D.dtype, device=D.device)
        out_I = torch.empty((world, nq, k), dtype=I.dtype, device=I.device)
    # flat (world*nq, k) view: the concatenation layout both NCCL and gloo accept; memory order is [rank][nq][k]
    dist.all_gather_into_tensor(out_D.view(world * nq, k), D.contiguous())
    dist.all_gather_into_tensor(out_I.view(world * nq, k), I.contiguous())
    return out_D, out_I


def sharded_search(local_search, merge, q, k, out_D=None, out_I=None):
    """local_search(q, k) -> (D, I) on this rank's shard (global ids); merge([R][nq][k] x2) -> (nq,k) x2"""
    D, I = local_search(q, k)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return D, I
    gD, gI = gather_topk(D, I, out_D, out_I)
    return merge(gD, gI)
