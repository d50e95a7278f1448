import time, torch, numpy as np, sys
sys.path.insert(0, '.')
from vector_line_quantization_b200 import index as vi, data, train, ops
dev = torch.device('cuda:0')
d, C, E, M = 128, 65536, 32, 16
gen = data.SyntheticGen('sift', d=d, kc=1 << 18, device=dev)
xt = torch.cat([gen.chunk(7000 + i, 1 << 20) for i in range(2)])
model = train.train_vlq(xt, C, E, M, 256, niter=2, pq_niter=4, exact_perm=False)
res = vi.StandardGpuResources(0)
h = vi.GpuIndexIVFPQ(res, d, C, M, 8, E, 256)
h.setCodebooks(*(model[k].cpu().numpy() for k in ("cent", "edge", "edge_d2", "lambda_cb", "pq")))
n = 1 << 21
h.reserveMemory(5 * n)
x = gen.chunk(5, 1 << 20); x = torch.cat([x, gen.chunk(6, 1 << 20)])
hx = torch.empty((n, d), dtype=torch.float32).pin_memory(); hx.copy_(x)
ids = torch.arange(n, dtype=torch.int64).pin_memory()
dx = x.clone(); dids = ids.to(dev)
def t(f, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-40s %.1f ms  %.1f Mvec/s" % (name, dt * 1e3, n / dt / 1e6), flush=True)
tmp = torch.empty_like(dx)
t(lambda: tmp.copy_(hx, non_blocking=True), "plain pinned H2D of the chunk (1.07 GB)")
t(lambda: tmp.copy_(hx, non_blocking=True), "plain pinned H2D of the chunk (1.07 GB)")
t(lambda: h.add_with_ids(dx, dids), "add_with_ids, device pointers (1st)")
t(lambda: h.add_with_ids(dx, dids), "add_with_ids, device pointers")
t(lambda: h.add_with_ids(hx, ids), "add_with_ids, pinned host (1st)")
t(lambda: h.add_with_ids(hx, ids), "add_with_ids, pinned host")
t(lambda: h.add_with_ids(hx, ids), "add_with_ids, pinned host")
t(lambda: h.search(x[:8].cpu().numpy(), 10), "first search (commit of 10 Mi entries)")
