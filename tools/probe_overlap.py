#!/usr/bin/env python
"""Probe: does running the coarse stage of query tile i+1 on a second stream while tile i is scanned shorten the step?
(C2 geometry on synthetic lists; kernels cannot share an SM with the tcgen05 sweep, so only tails can overlap.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from bench_scan import synthetic_lists
    from vector_line_quantization_b200 import data, ops

    dev = torch.device("cuda:0")
    C, d, E, P, W, k, M, nq = 65536, 128, 32, 64, 256, 100, 16, 10000
    tile = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cent = torch.from_numpy(data.sift_like(C, d=d, kc=1024, seed=1)).to(dev)
    q = torch.from_numpy(data.sift_like(nq, d=d, kc=1024, seed=2)).to(dev)
    pack = ops.CentPack(cent)
    edge, ed2 = ops.knn_graph(cent, E)
    lists, _, _ = synthetic_lists(ops, 10_000_000, C * E, M, dev)
    g = torch.Generator(device=dev).manual_seed(7)
    pq = torch.randn(M, 256, d // M, device=dev, generator=g)
    lcb = torch.rand(256, device=dev, generator=g)
    ed2f = ed2.reshape(-1)
    nb = ops.num_buckets(C)
    Db = [torch.empty((tile, C), dtype=torch.float32, device=dev) for _ in range(2)]
    bm = [torch.empty((tile, nb), dtype=torch.float32, device=dev) for _ in range(2)]
    ln = [tuple(torch.empty((tile, W), dtype=dt, device=dev) for dt in (torch.int32, torch.float32, torch.float32)) for _ in range(2)]
    outD = torch.empty((nq, k), dtype=torch.float32, device=dev)
    outI = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def coarse(i, s, e):
        m = e - s
        D = ops.l2_distances_tc(q[s:e], pack, out=Db[i][:m], bucket_min=bm[i][:m])
        ops.coarse_select_lines(D, bm[i][:m], C, P, edge, ed2, W, out=tuple(t[:m] for t in ln[i]))

    def scan(i, s, e):
        m = e - s
        l, t1, t6 = (t[:m] for t in ln[i])
        ops.scan_topk(q[s:e], pq, lcb, l, t1, t6, ed2f, lists, k, 1024, list_len_hint=4, out=(outD[s:e], outI[s:e]))

    tiles = [(s, min(nq, s + tile)) for s in range(0, nq, tile)]

    def serial():
        for i, (s, e) in enumerate(tiles):
            coarse(i & 1, s, e)
            scan(i & 1, s, e)

    def overlapped():
        cur = torch.cuda.current_stream()
        sa.wait_stream(cur)
        sb.wait_stream(cur)
        done_scan = [None, None]
        for i, (s, e) in enumerate(tiles):
            b = i & 1
            with torch.cuda.stream(sa):
                if done_scan[b] is not None:
                    sa.wait_event(done_scan[b])  # the line buffers of this slot are free again
                coarse(b, s, e)
                ev = torch.cuda.Event()
                ev.record(sa)
            with torch.cuda.stream(sb):
                sb.wait_event(ev)
                scan(b, s, e)
                done_scan[b] = torch.cuda.Event()
                done_scan[b].record(sb)
        cur.wait_stream(sa)
        cur.wait_stream(sb)

    res = {}
    for name, fn in (("serial", serial), ("overlapped", overlapped)):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = {"ms_best": min(ts), "ms_mean": sum(ts) / len(ts), "checksum": int(outI.clamp_min(0).sum()) & 0xffffffff}
    res["tile"] = tile
    print(json.dumps(res))


if __name__ == "__main__":
    main()
