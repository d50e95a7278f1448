#!/usr/bin/env python
"""HBM write-only / read-only / copy bandwidth (torch kernels, CUDA events, best of 10): the roofline of kernels that only
WRITE (the coarse distance tile of the query path) is the write-only figure, not the copy figure of MEASURED_PEAKS.json."""
import json

import torch

dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


af = a.view(torch.float32)
out = {
    "fill_GBs": n / best(lambda: af.fill_(1.0)) / 1e6,
    "copy_GBs_read_plus_write": 2 * n / best(lambda: b.copy_(a)) / 1e6,
    "read_sum_GBs": n / best(lambda: af.sum()) / 1e6,
}
print(json.dumps(out))
