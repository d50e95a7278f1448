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


def route_by_list(new_list, arrays, nlists, group=None):
    """List-range shards (the reference's own split, readDbFromFile(name, pronum, rank), gpu/GpuIndexIVFPQ.cu:2132-2163):
    rank r owns lists [r * nlists / R, (r + 1) * nlists / R).  Every rank has encoded its rows; this sends each encoded
    entry (list id + the per-entry arrays) to the owner of its list with one variable-size all-to-all per array.
    Entries without a list (-1) are dropped.  Returns (list ids, arrays) of the entries this rank now owns, ordered by
    source rank and, inside a source, by arrival -- i.e. by global row when the ranks hold consecutive id ranges."""
    world = dist.get_world_size(group)
    owner = torch.div(new_list.to(torch.int64) * world, nlists, rounding_mode="floor")
    owner = torch.where(new_list >= 0, owner, torch.full_like(owner, world))  # world = dropped
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world + 1)[:world]
    recv = torch.empty_like(counts)
    dist.all_to_all_single(recv, counts, group=group)
    send_sizes = counts.tolist()
    recv_sizes = recv.tolist()
    n_send, n_recv = sum(send_sizes), sum(recv_sizes)
    order = order[:n_send]

    def xchg(t):
        src = t[order].contiguous()
        out = torch.empty((n_recv,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(out, src, output_split_sizes=recv_sizes, input_split_sizes=send_sizes, group=group)
        return out

    return xchg(new_list), [xchg(t) for t in arrays]


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


class QuerySplitSearch:
    """Strong-scaling form of the sharded search (database split R ways, fixed query batch):

      1. the coarse stage (distances to the C centroids, top-P, line selection) does not touch the inverted lists, so
         rank r runs it for ITS 1/R of the queries only and writes (list, term1, term6) -- 12 W bytes per query -- into
         its peer-mapped symmetric buffer;
      2. one device barrier; every rank pulls the other ranks' line slices over NVLink (vlq_gather_peer_slices: one
         kernel of P2P loads, (R - 1) x nq/R x 12 W bytes) and scans ALL queries on its shard of the lists;
      3. one device barrier; rank r merges the R per-shard top-k lists of ITS query slice straight from the peers'
         buffers (vlq_merge_topk_peers with row offsets) -- the final (nq, k) result is distributed by query slice.

    Replicated work per rank is only the term-3 tables (16 KB per query).  Buffers are double-buffered per step like
    PeerExchange.  Results equal the single-index search exactly up to ties (the merge is exact)."""

    def __init__(self, nq, k, W, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.nq, self.k, self.W = nq, k, W
        self.q0 = [r * nq // self.world for r in range(self.world + 1)]  # query slice of rank r: [q0[r], q0[r+1])
        a256 = lambda v: (v + 255) // 256 * 256
        self.off_l = 0
        self.off_t1 = a256(self.off_l + nq * W * 4)
        self.off_t6 = a256(self.off_t1 + nq * W * 4)
        self.off_d = a256(self.off_t6 + nq * W * 4)
        self.off_i = a256(self.off_d + nq * k * 4)
        self.slot = a256(self.off_i + nq * k * 8)
        self.buf = symm.empty(2 * self.slot, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs_dev = self.hdl.buffer_ptrs_dev
        self.step = 0
        self.device = device
        self._peer = {}

    def _views(self, base_tensor, b):
        nq, k, W = self.nq, self.k, self.W
        v = lambda off, n, dt, shape: base_tensor[b + off:b + off + n].view(dt).view(shape)
        return (v(self.off_l, nq * W * 4, torch.int32, (nq, W)), v(self.off_t1, nq * W * 4, torch.float32, (nq, W)),
                v(self.off_t6, nq * W * 4, torch.float32, (nq, W)), v(self.off_d, nq * k * 4, torch.float32, (nq, k)),
                v(self.off_i, nq * k * 8, torch.int64, (nq, k)))

    def _peer_buf(self, r):
        if r not in self._peer:
            self._peer[r] = self.hdl.get_buffer(r, (2 * self.slot,), torch.uint8)
        return self._peer[r]

    def my_slice(self):
        return self.q0[self.rank], self.q0[self.rank + 1]

    def search(self, q, coarse_fn, scan_fn, out=None):
        """coarse_fn(q_slice, out=(lst, t1, t6)); scan_fn(q, (lst, t1, t6), out=(D, I)) on this rank's shard.
        Returns (D, I) of THIS rank's query slice [q0[rank], q0[rank + 1])."""
        from . import ops

        b = (self.step & 1) * self.slot
        self.step += 1
        lst, t1, t6, D, I = self._views(self.buf, b)
        s, e = self.my_slice()
        coarse_fn(q[s:e], out=(lst[s:e], t1[s:e], t6[s:e]))
        self.hdl.barrier(channel=0)  # every rank's line slice is written
        if (self.W * 4) % 16 == 0:  # one kernel of P2P loads pulls the other ranks' slices of the three arrays
            ops.gather_peer_slices(self.ptrs_dev, self.world, self.rank,
                                   [b + self.off_l, b + self.off_t1, b + self.off_t6], self.nq, self.W * 4)
        else:
            for r in range(self.world):
                if r == self.rank:
                    continue
                pl, p1, p6, _, _ = self._views(self._peer_buf(r), b)
                rs, re = self.q0[r], self.q0[r + 1]
                lst[rs:re].copy_(pl[rs:re], non_blocking=True)
                t1[rs:re].copy_(p1[rs:re], non_blocking=True)
                t6[rs:re].copy_(p6[rs:re], non_blocking=True)
        scan_fn(q, (lst, t1, t6), out=(D, I))
        self.hdl.barrier(channel=0)  # every shard's results are written
        n = e - s
        if out is None:
            out = (torch.empty((n, self.k), dtype=torch.float32, device=self.device),
                   torch.empty((n, self.k), dtype=torch.int64, device=self.device))
        if n > 0:
            ops.merge_topk_peers(self.ptrs_dev, b + self.off_d + s * self.k * 4, b + self.off_i + s * self.k * 8, self.world,
                                 n, self.k, out=out, device=self.device)
        return out


def sharded_search(local_search, merge, q, k, out_D=None, out_I=None):
    """local_search(q, k) -> (D, I) on this rank's shard (global ids); merge([R][nq][k] x2) -> (nq,k) x2"""
    D, I = local_search(q, k)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return D, I
    gD, gI = gather_topk(D, I, out_D, out_I)
    return merge(gD, gI)
