"""Device-resident operator layer: torch CUDA tensors in, torch CUDA tensors out, every op one C-ABI call.

torch is plumbing here (device memory + the current stream); all arithmetic happens in lib/libvlq_b200.so.
Each function names the C-ABI entry it calls; the reference interface that entry replaces is cited in
include/vlq_b200.h.
"""
from collections import namedtuple

import os

import torch

from . import _abi

Lists = namedtuple("Lists", "offsets codes lamq kappa ids")  # CSR inverted lists (see csrc/lists.cu)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def launch_count():
    return int(_abi.lib().vlq_launch_count())


def row_norms(x):
    x = _chk(x, torch.float32, "x")
    n, d = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    _abi.call("vlq_row_norms", _ptr(x), n, d, _ptr(out), _stream())
    return out


def l2_assign(x, cent, cnorm=None, add_xnorm=True, want_dist=True):
    """nearest centroid per row (a2): -> (ids int32 [n], dist f32 [n] or None)"""
    x = _chk(x, torch.float32, "x")
    cent = _chk(cent, torch.float32, "cent")
    n, d = x.shape
    if cnorm is None:
        cnorm = row_norms(cent)
    ids = torch.empty(n, dtype=torch.int32, device=x.device)
    dist = torch.empty(n, dtype=torch.float32, device=x.device) if want_dist else None
    _abi.call("vlq_l2_assign", _ptr(x), n, d, _ptr(cent), _ptr(cnorm), cent.shape[0], int(add_xnorm), _ptr(ids),
              _ptr(dist), _stream())
    return ids, dist


def l2_distances(x, cent, cnorm=None, out=None):
    """coarse matrix D = ||c||^2 - 2 x.c (a11)"""
    x = _chk(x, torch.float32, "x")
    cent = _chk(cent, torch.float32, "cent")
    n, d = x.shape
    C = cent.shape[0]
    if cnorm is None:
        cnorm = row_norms(cent)
    D = out if out is not None else torch.empty((n, C), dtype=torch.float32, device=x.device)
    _abi.call("vlq_l2_distances", _ptr(x), n, d, _ptr(cent), _ptr(cnorm), C, _ptr(D), D.stride(0), _stream())
    return D


def select_rows(D, k, row_add=None, cols=None):
    """exact top-k per row ascending (a11/a15): -> (val f32 [n][k], idx int32 [n][k])"""
    D = _chk(D, torch.float32, "D")
    n = D.shape[0]
    cols = D.shape[1] if cols is None else cols
    val = torch.empty((n, k), dtype=torch.float32, device=D.device)
    idx = torch.empty((n, k), dtype=torch.int32, device=D.device)
    _abi.call("vlq_select_rows", _ptr(D), n, cols, D.stride(0), k, _ptr(row_add), _ptr(val), _ptr(idx), _stream())
    return val, idx


def knn_graph(cent, E, cnorm=None):
    """centroid kNN graph (a4): -> (edge int32 [C][E], edge_d2 f32 [C][E])"""
    cent = _chk(cent, torch.float32, "cent")
    C, d = cent.shape
    if cnorm is None:
        cnorm = row_norms(cent)
    edge = torch.empty((C, E), dtype=torch.int32, device=cent.device)
    ed2 = torch.empty((C, E), dtype=torch.float32, device=cent.device)
    wsb = _abi.lib().vlq_knn_graph_workspace_bytes(C, E)
    ws = torch.empty(wsb, dtype=torch.uint8, device=cent.device)
    _abi.call("vlq_knn_graph", _ptr(cent), _ptr(cnorm), C, d, E, _ptr(edge), _ptr(ed2), _ptr(ws), wsb, _stream())
    return edge, ed2


Encoded = namedtuple("Encoded", "list lam lamq codes kappa residual")


def line_encode(x, assign, cent, edge, edge_d2, lambda_cb=None, pq=None, want_residual=False):
    """fused line stage (+ lambda quantiser + residual + PQ encode when lambda_cb/pq are given) (a5-a8)"""
    x = _chk(x, torch.float32, "x")
    assign = _chk(assign, torch.int32, "assign")
    cent = _chk(cent, torch.float32, "cent")
    edge = _chk(edge, torch.int32, "edge")
    edge_d2 = _chk(edge_d2, torch.float32, "edge_d2")
    n, d = x.shape
    E = edge.shape[1]
    dev = x.device
    out_list = torch.empty(n, dtype=torch.int32, device=dev)
    out_lam = torch.empty(n, dtype=torch.float32, device=dev)
    lamq = codes = kappa = resid = None
    M = nL = 0
    if lambda_cb is not None:
        lambda_cb = _chk(lambda_cb, torch.float32, "lambda_cb")
        pq = _chk(pq, torch.float32, "pq")
        M, nL = pq.shape[0], lambda_cb.shape[0]
        lamq = torch.empty(n, dtype=torch.uint8, device=dev)
        codes = torch.empty((n, M), dtype=torch.uint8, device=dev)
        kappa = torch.empty(n, dtype=torch.float32, device=dev)
        if want_residual:
            resid = torch.empty((n, d), dtype=torch.float32, device=dev)
    _abi.call("vlq_line_encode", _ptr(x), n, d, _ptr(assign), _ptr(cent), _ptr(edge), _ptr(edge_d2), E,
              _ptr(lambda_cb), nL, _ptr(pq), M, _ptr(out_list), _ptr(out_lam), _ptr(lamq), _ptr(codes), _ptr(kappa),
              _ptr(resid), _stream())
    return Encoded(out_list, out_lam, lamq, codes, kappa, resid)


def build_lists(nlists, M, new_list, new_codes, new_lamq, new_kappa, new_ids, old=None):
    """stable counting-sort append of new entries behind the existing CSR lists (a9) -> Lists"""
    dev = new_list.device
    n_new = new_list.shape[0]
    n_old = 0 if old is None else old.ids.shape[0]
    tot = n_old + n_new
    out = Lists(
        torch.empty(nlists + 1, dtype=torch.int64, device=dev),
        torch.empty((tot, M), dtype=torch.uint8, device=dev),
        torch.empty(tot, dtype=torch.uint8, device=dev),
        torch.empty(tot, dtype=torch.float32, device=dev),
        torch.empty(tot, dtype=torch.int64, device=dev),
    )
    wsb = _abi.lib().vlq_build_lists_workspace_bytes(n_new, nlists)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    o = old if n_old > 0 else Lists(None, None, None, None, None)
    _abi.call("vlq_build_lists", nlists, M, n_old, _ptr(o.offsets), _ptr(o.codes), _ptr(o.lamq), _ptr(o.kappa),
              _ptr(o.ids), n_new, _ptr(_chk(new_list, torch.int32, "new_list")),
              _ptr(_chk(new_codes, torch.uint8, "new_codes")), _ptr(_chk(new_lamq, torch.uint8, "new_lamq")),
              _ptr(_chk(new_kappa, torch.float32, "new_kappa")), _ptr(_chk(new_ids, torch.int64, "new_ids")),
              _ptr(out.offsets), _ptr(out.codes), _ptr(out.lamq), _ptr(out.kappa), _ptr(out.ids), _ptr(ws), wsb,
              _stream())
    return out


def rotate_codes(offsets, codes, inverse=False):
    """canonical list-major codes [n][M] -> the stored layout (or back with inverse=True): the entry at position pos of
    its list keeps stored[j] = code[(j + pos) mod M] (csrc/scan.cuh).  Plumbing for callers that assemble Lists by hand
    (tests, fixtures); vlq_build_lists produces the stored layout itself."""
    n, M = codes.shape
    lens = offsets[1:] - offsets[:-1]
    live = int(offsets[-1])  # rows beyond the last list (skipped vectors) are left as they are
    pos = torch.zeros(n, dtype=torch.int64, device=codes.device)
    pos[:live] = torch.arange(live, device=codes.device) - torch.repeat_interleave(offsets[:-1], lens)
    j = torch.arange(M, device=codes.device).unsqueeze(0)
    idx = (j + (-pos if inverse else pos).unsqueeze(1)) % M
    return torch.gather(codes, 1, idx).contiguous()


def select_lines(D, coarse_ids, edge, edge_d2, W, out=None):
    """query-time line selection (a12): -> (list int32 [nq][W], term1, term6 f32 [nq][W])"""
    D = _chk(D, torch.float32, "D")
    coarse_ids = _chk(coarse_ids, torch.int32, "coarse_ids")
    nq, P = coarse_ids.shape
    E = edge.shape[1]
    dev = D.device
    if out is not None:
        lst, t1, t6 = out
        assert lst.shape == (nq, W) and lst.is_contiguous() and t1.is_contiguous() and t6.is_contiguous()
    else:
        lst = torch.empty((nq, W), dtype=torch.int32, device=dev)
        t1 = torch.empty((nq, W), dtype=torch.float32, device=dev)
        t6 = torch.empty((nq, W), dtype=torch.float32, device=dev)
    _abi.call("vlq_select_lines", _ptr(D), nq, D.stride(0), _ptr(coarse_ids), P, _ptr(edge), _ptr(edge_d2), E, W,
              _ptr(lst), _ptr(t1), _ptr(t6), _stream())
    return lst, t1, t6


_scan_ws = {}


def scan_topk(q, pq, lambda_cb, line_list, term1, term6, edge_d2, lists, k, cap=1024, use_workspace=True,
              list_len_hint=None, out=None):
    """ADC scan of the selected lists fused with exact top-k (a13-a15): -> (D f32 [nq][k], I int64 [nq][k])"""
    q = _chk(q, torch.float32, "q")
    pq = _chk(pq, torch.float32, "pq")
    nq, d = q.shape
    M = pq.shape[0]
    W = line_list.shape[1]
    if out is not None:  # contiguous (nq, k) f32 / int64 destinations (e.g. row slices of the caller's result arrays)
        outD, outI = _chk(out[0], torch.float32, "out D"), _chk(out[1], torch.int64, "out I")
        assert outD.shape == (nq, k) and outI.shape == (nq, k)
    else:
        outD = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        outI = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    hint = list_len_hint if list_len_hint is not None else lists.ids.shape[0] // max(1, lists.offsets.shape[0] - 1)
    ws = None
    wsb = 0
    if use_workspace:  # term-3 tables of the batch (grow-only cache per device)
        wsb = _abi.lib().vlq_scan_topk_workspace_bytes(nq, M)
        ws = _scan_ws.get(q.device)
        if ws is None or ws.numel() < wsb:
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=q.device)
            _scan_ws[q.device] = ws
    _abi.call("vlq_scan_topk", _ptr(q), nq, d, _ptr(pq), M, _ptr(lambda_cb), lambda_cb.shape[0],
              _ptr(_chk(line_list, torch.int32, "line_list")), _ptr(term1), _ptr(term6), _ptr(edge_d2), W,
              _ptr(lists.offsets), _ptr(lists.codes), _ptr(lists.lamq), _ptr(lists.kappa), _ptr(lists.ids), k, cap,
              int(hint), _ptr(outD), _ptr(outI), _ptr(ws), wsb, _stream())
    return outD, outI


def merge_topk(D, I):
    """shard merge (a16): D, I are [R][nq][k] -> (nq,k)"""
    D = _chk(D, torch.float32, "D")
    I = _chk(I, torch.int64, "I")
    R, nq, k = D.shape
    outD = torch.empty((nq, k), dtype=torch.float32, device=D.device)
    outI = torch.empty((nq, k), dtype=torch.int64, device=D.device)
    _abi.call("vlq_merge_topk", _ptr(D), _ptr(I), R, nq, k, _ptr(outD), _ptr(outI), _stream())
    return outD, outI


def merge_topk_peers(peer_ptrs_dev, d_off, i_off, R, nq, k, out=None, device=None):
    """shard merge with the gather fused in (a16): peer_ptrs_dev = device address of an array of R buffer pointers
    (symmetric memory), each buffer holding [D f32 (nq,k)] at d_off and [I int64 (nq,k)] at i_off"""
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=device), torch.empty((nq, k), dtype=torch.int64, device=device))
    _abi.call("vlq_merge_topk_peers", int(peer_ptrs_dev), d_off, i_off, R, nq, k, _ptr(out[0]), _ptr(out[1]), _stream())
    return out


def gather_peer_slices(peer_ptrs_dev, R, rank, arr_offsets, nq, row_bytes):
    """query-split sharding: pull the other ranks' row slices of the arrays at arr_offsets (bytes inside the symmetric
    buffers) into this rank's buffer, one launch of P2P loads"""
    import ctypes

    offs = (ctypes.c_int64 * len(arr_offsets))(*[int(o) for o in arr_offsets])
    _abi.call("vlq_gather_peer_slices", int(peer_ptrs_dev), R, rank, ctypes.addressof(offs), len(arr_offsets), nq,
              row_bytes, _stream())


def km_update(x, assign, k):
    """k-means mean step, deterministic row order (f1): -> (centroids f32 [k][d], counts int32 [k])"""
    x = _chk(x, torch.float32, "x")
    assign = _chk(assign, torch.int32, "assign")
    n, d = x.shape
    cent = torch.empty((k, d), dtype=torch.float32, device=x.device)
    counts = torch.empty(k, dtype=torch.int32, device=x.device)
    wsb = _abi.lib().vlq_km_update_workspace_bytes(n, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    _abi.call("vlq_km_update", _ptr(x), n, d, _ptr(assign), k, _ptr(cent), _ptr(counts), _ptr(ws), wsb, _stream())
    return cent, counts


def _tile_rows(nq, tile):
    """equal query tiles of at most `tile` rows, multiples of 256 (GpuIndexIVFPQ::search makes the same split)"""
    if nq <= tile:
        return max(nq, 1)
    nt = -(-nq // tile)
    return min(tile, (-(-nq // nt) + 255) // 256 * 256)


def coarse_lines(q, cent, cnorm, edge, edge_d2, P, W, tile=5120, pack=None, out=None):
    """First half of the query path (a11 + a12) for a batch of queries: -> (list int32, term1, term6 f32), each [nq][W].
    This half does not touch the inverted lists, so shards can split the QUERIES for it (sharding.QuerySplitSearch)."""
    nq = q.shape[0]
    C = cent.shape[0]
    P = min(P, C)
    if out is None:
        out = (torch.empty((nq, W), dtype=torch.int32, device=q.device), torch.empty((nq, W), dtype=torch.float32, device=q.device),
               torch.empty((nq, W), dtype=torch.float32, device=q.device))
    if nq == 0:
        return out
    tile = _tile_rows(nq, tile)
    stage = CoarseStage(cent, cnorm, edge, edge_d2, P, W, tile, pack)
    for s in range(0, nq, tile):
        e = min(nq, s + tile)
        stage.run(q[s:e], out=tuple(t[s:e] for t in out))
    return out


class CoarseStage:
    """a11 + a12 for one query tile -- the dispatch GpuIndexIVFPQ::coarseLines_ makes.  With a CentPack and a small
    nprobe (vlq_coarse_exact_preferred) the distance matrix is never written: the tcgen05 sweep emits the bucket minima
    only and vlq_coarse_select_lines_exact re-evaluates the columns it needs; otherwise vlq_l2_distances_tc +
    vlq_coarse_select_lines (VLQ_COARSE_MATRIX=1 forces this route, VLQ_COARSE_EXACT=1 the other wherever it is
    supported).  Without a pack: fp32 matrix + select_rows + select_lines."""

    def __init__(self, cent, cnorm, edge, edge_d2, P, W, tile, pack=None):
        self.cent, self.cnorm, self.edge, self.edge_d2, self.pack = cent, cnorm, edge, edge_d2, pack
        self.C, self.d = cent.shape
        self.P, self.W = min(P, self.C), W
        dev = cent.device
        fn = "vlq_coarse_exact_supported" if os.environ.get("VLQ_COARSE_EXACT") else "vlq_coarse_exact_preferred"
        self.exact = (pack is not None and os.environ.get("VLQ_COARSE_MATRIX") is None and
                      bool(getattr(_abi.lib(), fn)(self.d, self.C, self.P, edge.shape[1], W)))
        self.Dbuf = None if self.exact else torch.empty((tile, self.C), dtype=torch.float32, device=dev)
        self.bbuf = torch.empty((tile, num_buckets(self.C)), dtype=torch.float32, device=dev) if pack is not None else None

    def run(self, qt, out=None, want_coarse=False):
        m = qt.shape[0]
        if self.exact:
            l2_bucket_min_tc(qt, self.pack, self.bbuf[:m])
            return coarse_select_lines_exact(qt, self.cent, self.cnorm, self.bbuf[:m], self.P, self.edge, self.edge_d2,
                                             self.W, want_coarse=want_coarse, out=out)
        if self.pack is not None:
            D = l2_distances_tc(qt, self.pack, out=self.Dbuf[:m], bucket_min=self.bbuf[:m])
            return coarse_select_lines(D, self.bbuf[:m], self.C, self.P, self.edge, self.edge_d2, self.W,
                                       want_coarse=want_coarse, out=out)
        D = l2_distances(qt, self.cent, self.cnorm, out=self.Dbuf[:m])
        _, cid = select_rows(D, self.P)
        r = select_lines(D, cid, self.edge, self.edge_d2, self.W, out=out)
        return (*r, cid) if want_coarse else r


def scan_lines(q, pq, lambda_cb, lines, edge_d2, lists, k, cap=1024, tile=5120, out=None, list_len_hint=None):
    """Second half of the query path (a13-a15): scan the selected lines of every query on THIS shard's lists"""
    nq = q.shape[0]
    lst, t1, t6 = lines
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=q.device), torch.empty((nq, k), dtype=torch.int64, device=q.device))
    ed2_flat = edge_d2.reshape(-1)
    tile = _tile_rows(nq, tile)
    for s in range(0, nq, tile):
        e = min(nq, s + tile)
        scan_topk(q[s:e], pq, lambda_cb, lst[s:e], t1[s:e], t6[s:e], ed2_flat, lists, k, cap, out=(out[0][s:e], out[1][s:e]),
                  list_len_hint=list_len_hint)
    return out


def search(q, cent, cnorm, edge, edge_d2, lambda_cb, pq, lists, P, W, k, cap=1024, tile=5120, pack=None, out=None,
           list_len_hint=None):
    """Full query path on resident tensors (a11-a15), tiled over queries (at most 5120 per tile: 1.25 GiB of coarse distances).
    pack (a CentPack) routes the coarse distances through the tcgen05 kernel."""
    nq = q.shape[0]
    C = cent.shape[0]
    P = min(P, C)
    if out is not None:
        outD, outI = out
    else:
        outD = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        outI = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    tile = _tile_rows(nq, tile)
    stage = CoarseStage(cent, cnorm, edge, edge_d2, P, W, tile, pack)
    ed2_flat = edge_d2.reshape(-1)
    for s in range(0, nq, tile):
        e = min(nq, s + tile)
        qt = q[s:e]
        lst, t1, t6 = stage.run(qt)
        scan_topk(qt, pq, lambda_cb, lst, t1, t6, ed2_flat, lists, k, cap, out=(outD[s:e], outI[s:e]),
                  list_len_hint=list_len_hint)
    return outD, outI


# ------------------------------------------------------------------------------------------------ tensor-core coarse path
class CentPack:
    """Pre-packed centroids for the tcgen05 kernels (vlq_tc_pack_centroids): fp16 hi/lo tiles + padded ||c||^2."""

    def __init__(self, cent, cnorm=None, scale=None):
        cent = _chk(cent, torch.float32, "cent")
        self.C, self.d = cent.shape
        if not _abi.lib().vlq_tc_supported(self.d, self.C):
            raise ValueError("tensor-core path needs d %% 32 == 0 and 32 <= d <= 128 (got d=%d)" % self.d)
        self.cnorm = row_norms(cent) if cnorm is None else cnorm
        if scale is None:
            import math

            mx = float(cent.abs().max())
            scale = 2.0 ** (9 - math.ceil(math.log2(mx))) if mx > 0 else 1.0
        self.scale = float(scale)
        nbytes = _abi.lib().vlq_tc_cent_pack_bytes(self.C, self.d)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=cent.device)
        _abi.call("vlq_tc_pack_centroids", _ptr(cent), _ptr(self.cnorm), self.C, self.d, self.scale, _ptr(self.buf),
                  _stream())
        self._ws = None

    def workspace(self, n):
        need = _abi.lib().vlq_l2_tc_workspace_bytes(n, self.d, self.C)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.buf.device)
        return self._ws


def l2_assign_tc(x, pack, add_xnorm=True, want_dist=True):
    """nearest centroid per row on the tensor cores (a2): -> (ids int32 [n], dist f32 [n] or None)"""
    x = _chk(x, torch.float32, "x")
    n, d = x.shape
    ids = torch.empty(n, dtype=torch.int32, device=x.device)
    dist = torch.empty(n, dtype=torch.float32, device=x.device) if want_dist else None
    ws = pack.workspace(n)
    _abi.call("vlq_l2_assign_tc", _ptr(x), n, d, _ptr(pack.buf), pack.scale, pack.C,
              1 if add_xnorm else 0, _ptr(ids),
              _ptr(dist), _ptr(ws), ws.numel(), _stream())
    return ids, dist


def l2_distances_tc(x, pack, out=None, bucket_min=None):
    """coarse matrix D = ||c||^2 - 2 x.c on the tensor cores (a11); bucket_min ([n][num_buckets] f32) optionally
    receives the minimum of every 32-column bucket (input of coarse_select_lines)"""
    x = _chk(x, torch.float32, "x")
    n, d = x.shape
    D = out if out is not None else torch.empty((n, pack.C), dtype=torch.float32, device=x.device)
    ws = pack.workspace(n)
    _abi.call("vlq_l2_distances_tc", _ptr(x), n, d, _ptr(pack.buf), pack.scale, pack.C, _ptr(D), D.stride(0),
              _ptr(bucket_min), _ptr(ws), ws.numel(), _stream())
    return D


def l2_bucket_min_tc(x, pack, bucket_min):
    """the tcgen05 sweep without the distance matrix: only the minima of the 32-column buckets are written"""
    x = _chk(x, torch.float32, "x")
    n, d = x.shape
    assert bucket_min.shape == (n, num_buckets(pack.C)) and bucket_min.is_contiguous()
    ws = pack.workspace(n)
    _abi.call("vlq_l2_bucket_min_tc", _ptr(x), n, d, _ptr(pack.buf), pack.scale, pack.C, _ptr(bucket_min), _ptr(ws),
              ws.numel(), _stream())
    return bucket_min


def coarse_select_lines_exact(q, cent, cnorm, bucket_min, P, edge, edge_d2, W, want_coarse=False, out=None):
    """a11 + a12 without a distance matrix (vlq_coarse_select_lines_exact); out = (list, term1, term6) destinations"""
    q = _chk(q, torch.float32, "q")
    nq, d = q.shape
    C, E = cent.shape[0], edge.shape[1]
    dev = q.device
    if out is not None:
        lst, t1, t6 = out
        assert lst.shape == (nq, W) and lst.is_contiguous() and t1.is_contiguous() and t6.is_contiguous()
    else:
        lst = torch.empty((nq, W), dtype=torch.int32, device=dev)
        t1 = torch.empty((nq, W), dtype=torch.float32, device=dev)
        t6 = torch.empty((nq, W), dtype=torch.float32, device=dev)
    cid = torch.empty((nq, P), dtype=torch.int32, device=dev) if want_coarse else None
    _abi.call("vlq_coarse_select_lines_exact", _ptr(q), nq, d, _ptr(cent), _ptr(cnorm), _ptr(bucket_min),
              bucket_min.shape[1], C, P, _ptr(edge), _ptr(edge_d2), E, W, _ptr(cid), _ptr(lst), _ptr(t1), _ptr(t6),
              _stream())
    return (lst, t1, t6, cid) if want_coarse else (lst, t1, t6)


def num_buckets(C):
    return int(_abi.lib().vlq_tc_num_buckets(C))


def coarse_select_lines(D, bucket_min, C, P, edge, edge_d2, W, want_coarse=False, out=None):
    """fused top-P (through the bucket minima) + line selection (a11 + a12); out = (list, term1, term6) destinations"""
    nq = D.shape[0]
    E = edge.shape[1]
    dev = D.device
    if out is not None:
        lst, t1, t6 = out
        assert lst.shape == (nq, W) and lst.is_contiguous() and t1.is_contiguous() and t6.is_contiguous()
    else:
        lst = torch.empty((nq, W), dtype=torch.int32, device=dev)
        t1 = torch.empty((nq, W), dtype=torch.float32, device=dev)
        t6 = torch.empty((nq, W), dtype=torch.float32, device=dev)
    cid = torch.empty((nq, P), dtype=torch.int32, device=dev) if want_coarse else None
    _abi.call("vlq_coarse_select_lines", _ptr(D), nq, D.stride(0), _ptr(bucket_min), bucket_min.shape[1], C, P,
              _ptr(edge), _ptr(edge_d2), E, W, _ptr(cid), _ptr(lst), _ptr(t1), _ptr(t6), _stream())
    return (lst, t1, t6, cid) if want_coarse else (lst, t1, t6)
