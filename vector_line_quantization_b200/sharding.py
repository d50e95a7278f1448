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


class PeerExchange:
    """The exchange step without a collective: every rank's scan writes its (nq,k) results into a buffer of peer-mapped
    symmetric memory (one allocation per rank, every rank holds the device pointers of all of them), and the merge kernel
    of each rank reads the R shards' results straight from the peers over NVLink (vlq_merge_topk_peers: gather + merge
    in ONE kernel of P2P loads).  Cross-GPU ordering is one device-side barrier per step, enqueued on the stream behind
    the scan; the buffers are double-buffered so that no second barrier is needed before the next step overwrites them
    (a rank reaches the barrier of step s+1 only after its merge of step s has run, and slot s%2 is rewritten in step
    s+2, behind that barrier).  Results are bit-identical to gather_topk + merge_topk."""

    def __init__(self, nq, k, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.world = dist.get_world_size()
        self.nq, self.k = nq, k
        self.i_off = (nq * k * 4 + 15) // 16 * 16
        self.slot = (self.i_off + nq * k * 8 + 255) // 256 * 256
        self.buf = symm.empty(2 * self.slot, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs_dev = self.hdl.buffer_ptrs_dev
        self.step = 0

    def local_out(self):
        """(D, I) destinations of this step's local search: views into the symmetric buffer"""
        b = (self.step & 1) * self.slot
        n = self.nq * self.k
        D = self.buf[b:b + 4 * n].view(torch.float32).view(self.nq, self.k)
        I = self.buf[b + self.i_off:b + self.i_off + 8 * n].view(torch.int64).view(self.nq, self.k)
        return D, I

    def merge(self, out=None):
        from . import ops

        self.hdl.barrier(channel=0)  # stream-ordered behind this rank's scan; returns once every peer has got there
        b = (self.step & 1) * self.slot
        self.step += 1
        return ops.merge_topk_peers(self.ptrs_dev, b, b + self.i_off, self.world, self.nq, self.k, out=out,
                                    device=self.buf.device)


def sharded_search(local_search, merge, q, k, out_D=None, out_I=None):
    """local_search(q, k) -> (D, I) on this rank's shard (global ids); merge([R][nq][k] x2) -> (nq,k) x2"""
    D, I = local_search(q, k)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return D, I
    gD, gI = gather_topk(D, I, out_D, out_I)
    return merge(gD, gI)
