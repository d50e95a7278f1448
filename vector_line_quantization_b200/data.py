"""Deterministic synthetic SIFT-/DEEP-shaped data: no real datasets are available offline.

SIFT-shape (d=128): Kc cluster centres mu_j ~ U{0..127}^d; vector = clip(round(mu_j + N(0, sigma^2)), 0, 255).
DEEP-shape (d=96):  normalize(mu_j/||mu_j|| + N(0, sigma^2)),  mu_j ~ N(0,1).
numpy versions are used by the tests (small, bit-reproducible); torch versions generate on the GPU for bench.py.
"""
import numpy as np


def sift_like(n, d=128, kc=4096, sigma=24.0, seed=1, centre_seed=1234, dtype=np.float32):
    crng = np.random.RandomState(centre_seed)
    centres = crng.randint(0, 128, size=(kc, d)).astype(np.float32)
    rng = np.random.RandomState(seed)
    j = rng.randint(0, kc, size=n)
    x = centres[j] + rng.normal(0.0, sigma, size=(n, d)).astype(np.float32)
    x = np.clip(np.rint(x), 0, 255)
    return np.ascontiguousarray(x.astype(dtype))


def deep_like(n, d=96, kc=4096, sigma=0.08, seed=1, centre_seed=1234):
    crng = np.random.RandomState(centre_seed)
    centres = crng.normal(size=(kc, d)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    rng = np.random.RandomState(seed)
    j = rng.randint(0, kc, size=n)
    x = centres[j] + rng.normal(0.0, sigma, size=(n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x.astype(np.float32))


def sift_like_torch(n, d=128, kc=1 << 18, sigma=24.0, seed=1, centre_seed=1234, device="cuda", chunk=1 << 20,
                    out_dtype=None):
    """GPU generator for bench.py (same distribution, different stream than the numpy version)."""
    import torch

    out_dtype = out_dtype or torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(centre_seed)
    centres = torch.randint(0, 128, (kc, d), generator=g, device=device, dtype=torch.int32).to(torch.float32)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=out_dtype, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        j = torch.randint(0, kc, (e - s,), generator=g, device=device)
        x = centres[j] + sigma * torch.randn((e - s, d), generator=g, device=device)
        out[s:e] = x.round_().clamp_(0, 255).to(out_dtype)
    return out


class SyntheticGen:
    """Chunk-wise deterministic generator on the GPU (bench.py streams databases that do not fit in one tensor).
    shape "sift": uint8-valued, d=128-like mixture; shape "deep": unit-norm float vectors."""

    def __init__(self, shape="sift", d=128, kc=1 << 18, sigma=None, centre_seed=1234, device="cuda"):
        import torch

        self.shape, self.d, self.kc, self.device = shape, d, kc, device
        g = torch.Generator(device=device)
        g.manual_seed(centre_seed)
        if shape == "sift":
            self.sigma = 24.0 if sigma is None else sigma
            self.centres = torch.randint(0, 128, (kc, d), generator=g, device=device, dtype=torch.int32).to(torch.float32)
        else:
            self.sigma = 0.08 if sigma is None else sigma
            c = torch.randn((kc, d), generator=g, device=device)
            self.centres = c / c.norm(dim=1, keepdim=True)

    def chunk(self, seed, n, dtype=None):
        import torch

        g = torch.Generator(device=self.device)
        g.manual_seed(int(seed))
        j = torch.randint(0, self.kc, (n,), generator=g, device=self.device)
        x = self.centres[j] + self.sigma * torch.randn((n, self.d), generator=g, device=self.device)
        if self.shape == "sift":
            x = x.round_().clamp_(0, 255)
            return x.to(dtype) if dtype is not None else x
        return x / x.norm(dim=1, keepdim=True)


def recall_at(I, gt, r):
    """fraction of queries whose true 1-NN (gt[i]) is among the first r returned labels
    (gpu/test/sift1b_query.cpp:334-347, tests/demo_sift1M.cpp:233-246)"""
    I = np.asarray(I)
    gt = np.asarray(gt).reshape(-1, 1)
    return float((I[:, :r] == gt).any(axis=1).mean())
