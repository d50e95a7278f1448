"""micro-benchmark of the tcgen05 coarse kernels (CUDA events, L2-flushed): python tools/bench_tc.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vector_line_quantization_b200 import ops, data

dev = torch.device("cuda:0")
C, d = 65536, 128
cent = data.sift_like_torch(C, d=d, kc=1 << 16, seed=5, device=dev) + 0.37
pack = ops.CentPack(cent)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=5):
    fn(); fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

for exact in (True, False):
    xq = data.sift_like_torch(4096, d=d, kc=1 << 16, seed=6, device=dev)
    xa = data.sift_like_torch(1 << 20, d=d, kc=1 << 16, seed=7, device=dev)
    if not exact:
        xq += 0.123; xa += 0.123   # lo parts non-zero -> 3 passes
    D = torch.empty((4096, C), device=dev)
    bm = torch.empty((4096, ops.num_buckets(C)), device=dev)
    t = timeit(lambda: ops.l2_distances_tc(xq, pack, out=D, bucket_min=bm))
    t2 = timeit(lambda: ops.l2_distances_tc(xq, pack, out=D))
    ta = timeit(lambda: ops.l2_assign_tc(xa, pack, want_dist=False))
    passes = 2 if exact else 3
    print("passes=%d  distances(4096 q) %.3f ms (no bmin %.3f ms) -> %.1f TFLOP/s alg | assign(1M) %.2f ms -> %.1f Mvec/s, %.1f TFLOP/s alg"
          % (passes, t[0], t2[0], 4096 * 2.0 * C * d / t[0] / 1e9, ta[0], (1 << 20) / ta[0] / 1e3, (1 << 20) * 2.0 * C * d / ta[0] / 1e9))
