#!/usr/bin/env python
"""Coarse stage (a11 + a12) micro-benchmark: the matrix path (vlq_l2_distances_tc + vlq_coarse_select_lines) against the
matrix-free path (vlq_l2_bucket_min_tc + vlq_coarse_select_lines_exact) over nprobe, one 4096-query tile per launch,
CUDA events, L2 flushed between repetitions.

  python tools/bench_coarse.py [--c 65536] [--d 128] [--e 32] [--nq 8192]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c", type=int, default=65536)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--e", type=int, default=32)
    ap.add_argument("--nq", type=int, default=8192)
    ap.add_argument("--tile", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    from vector_line_quantization_b200 import data, ops

    dev = torch.device("cuda:0")
    C, d, E = a.c, a.d, a.e
    cent = torch.from_numpy(data.sift_like(C, d=d, kc=1024, seed=1)).to(dev)
    q = torch.from_numpy(data.sift_like(a.nq, d=d, kc=1024, seed=2)).to(dev)
    pack = ops.CentPack(cent)
    edge, ed2 = ops.knn_graph(cent, E)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    Dbuf = torch.empty((a.tile, C), dtype=torch.float32, device=dev)
    bm = torch.empty((a.tile, ops.num_buckets(C)), dtype=torch.float32, device=dev)
    out = []
    for P in (1, 2, 4, 8, 16, 32, 64):
        W = min(1024, 4 * P)

        def run(exact):
            evs = []
            for s in range(0, a.nq, a.tile):
                qt = q[s:s + a.tile]
                m = qt.shape[0]
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                if exact:
                    ops.l2_bucket_min_tc(qt, pack, bm[:m])
                    e1.record()
                    ops.coarse_select_lines_exact(qt, cent, pack.cnorm, bm[:m], P, edge, ed2, W)
                else:
                    D = ops.l2_distances_tc(qt, pack, out=Dbuf[:m], bucket_min=bm[:m])
                    e1.record()
                    ops.coarse_select_lines(D, bm[:m], C, P, edge, ed2, W)
                e2.record()
                evs.append((e0, e1, e2))
            torch.cuda.synchronize()
            return sum(x.elapsed_time(y) for x, y, _ in evs), sum(y.elapsed_time(z) for _, y, z in evs)

        row = {"P": P, "W": W}
        for exact in (False, True):
            run(exact)
            best = None
            for _ in range(a.steps):
                flush.fill_(1)
                g, s_ = run(exact)
                if best is None or g + s_ < sum(best):
                    best = (g, s_)
            key = "exact" if exact else "matrix"
            row[key + "_gemm_ms"], row[key + "_select_ms"] = best
            row[key + "_ms"] = sum(best)
        row["exact_l2_KB_per_query"] = (32 * P + P * E) * d * 4 / 1024
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
