#!/usr/bin/env python
"""Scan-stage micro-benchmark at a chosen list density on SYNTHETIC lists (random codes / lambda bytes / kappa, no
encode): times vlq_scan_topk alone, tile by tile, with CUDA events, and reports algorithmic bytes / time against the HBM
peak ((M + 1) bytes per scanned entry + 8 k bytes per query).

  python tools/bench_scan.py [--entries 1e9] [--nlists 2097152] [--m 16] [--nq 10000] [--w1 256] [--k 100]
  VLQ_SCAN_KERNEL=skew|long python tools/bench_scan.py ...   # force the register-pipelined warp-autonomous kernel /
                                                             # the long-list kernel (default: chosen by list length)
  VLQ_SCAN_LOOK=n VLQ_SCAN_LPT=0|1                           # tuning knobs of the long-list kernel (scan_long.cu)

List lengths are Poisson around entries/nlists with a Gamma(4) spread and queries pick lists size-biased (dense regions
attract both vectors and queries, as in the real index: 193 k scanned entries per query at 1 B vs 256 x 477 = 122 k).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_lists(ops, n_target, nlists, M, dev, seed=1, keep_frac=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    mean = n_target / nlists
    spread = torch.distributions.Gamma(4.0, 4.0).sample((nlists,)).to(dev)
    lens_full = torch.poisson(spread * mean, generator=g).to(torch.int64)
    lens = lens_full.clone()
    if keep_frac < 1.0:  # a list-range shard: this rank holds the first keep_frac of the lists, the others are empty here
        lens[int(nlists * keep_frac):] = 0
    off = torch.zeros(nlists + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    n = int(off[-1])
    pad = 64
    codes = torch.empty((n + pad, M), dtype=torch.uint8, device=dev)
    lamq = torch.empty(n + pad, dtype=torch.uint8, device=dev)
    kappa = torch.empty(n + pad, dtype=torch.float32, device=dev)
    step = 1 << 26
    for s in range(0, n + pad, step):
        e = min(n + pad, s + step)
        codes[s:e] = torch.randint(0, 256, (e - s, M), dtype=torch.uint8, device=dev, generator=g)
        lamq[s:e] = torch.randint(0, 256, (e - s,), dtype=torch.uint8, device=dev, generator=g)
        kappa[s:e] = torch.randn(e - s, device=dev, generator=g) * 100.0
    ids = torch.arange(n, dtype=torch.int64, device=dev)
    return ops.Lists(off, codes[:n], lamq[:n], kappa[:n], ids), lens, lens_full


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entries", type=float, default=1e9)
    ap.add_argument("--nlists", type=int, default=1 << 21)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--w1", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--tile", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--hint", type=int, default=-1)
    ap.add_argument("--keep-frac", type=float, default=1.0,
                    help="emulate one rank of R list-range shards: only this fraction of the lists is populated, the "
                         "queries still select lines over the whole index (1/R of each query's lines are non-empty)")
    a = ap.parse_args()
    from vector_line_quantization_b200 import _abi, ops

    _abi.lib()
    dev = torch.device("cuda:0")
    M, W, k, nq = a.m, a.w1, a.k, a.nq
    lists, lens, lens_full = synthetic_lists(ops, int(a.entries), a.nlists, M, dev, keep_frac=a.keep_frac)
    g = torch.Generator(device=dev).manual_seed(7)
    line = torch.multinomial(lens_full.float() + 1e-3, nq * W, replacement=True, generator=g).reshape(nq, W).to(torch.int32)
    q = torch.randn(nq, a.d, device=dev, generator=g)
    pq = torch.randn(M, 256, a.d // M, device=dev, generator=g)
    lcb = torch.rand(256, device=dev, generator=g)
    t1 = torch.rand(nq, W, device=dev, generator=g) * 10
    t6 = torch.randn(nq, W, device=dev, generator=g)
    ed2 = torch.rand(a.nlists, device=dev, generator=g) * 4 + 0.5
    scanned = float(lens.clamp_max(1024)[line.to(torch.int64)].sum(dim=1).float().mean())
    hint = a.hint if a.hint >= 0 else int(a.entries / a.nlists)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    outD = torch.empty((nq, k), dtype=torch.float32, device=dev)
    outI = torch.empty((nq, k), dtype=torch.int64, device=dev)

    def run():
        evs = []
        for s in range(0, nq, a.tile):
            e = min(nq, s + a.tile)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.scan_topk(q[s:e], pq, lcb, line[s:e], t1[s:e], t6[s:e], ed2, lists, k, 1024, list_len_hint=hint,
                          out=(outD[s:e], outI[s:e]))
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(x.elapsed_time(y) for x, y in evs)

    run()
    run()
    ms = []
    for _ in range(a.steps):
        flush.fill_(1)
        ms.append(run())
    best, mean = min(ms), sum(ms) / len(ms)
    peak = 6556.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    alg = nq * (scanned * (M + 1) + 8 * k)
    streamed = nq * scanned * (M + 5)
    chk = int(outI.clamp_min(0).sum()) & 0xffffffff
    print(json.dumps({
        "kernel": os.environ.get("VLQ_SCAN_KERNEL", "auto"),
        "entries": int(lists.ids.shape[0]), "avg_len": int(a.entries) / a.nlists, "keep_frac": a.keep_frac, "M": M, "nq": nq, "w1": W, "k": k,
        "scanned_per_query": scanned, "ms_mean": mean, "ms_best": best,
        "algorithmic_GBs": alg / (mean * 1e-3) / 1e9, "streamed_GBs": streamed / (mean * 1e-3) / 1e9,
        "frac_of_hbm_peak": alg / (mean * 1e-3) / 1e9 / peak, "peak": peak, "checksum": chk}))


if __name__ == "__main__":
    main()
