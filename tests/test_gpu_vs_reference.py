"""CUDA path against the UNMODIFIED reference library itself (oracle/_ref/libfaiss_ref.so, the reference's CPU
IndexIVFPQ), not just against the oracle restatement.

With a single lambda level equal to 0 the VLQ anchor (1-l)c + l s is the centroid c itself, so a VLQ index whose search
keeps ALL lines of the probed centroids (w1 = nprobe * nedge) is exactly IndexIVFPQ with nprobe probes: same coarse
assignment, same residual PQ codes, same ADC distances (up to ||q||^2, which the GPU search path omits like the
reference's GPU path, gpu/impl/Distance.cu:287-290).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libfaiss_ref.so not built")
    oracle.ref()
    return oracle


@pytest.mark.parametrize("d,C,E,M,P", [(128, 64, 16, 16, 8), (64, 32, 8, 8, 6)])
def test_gpu_vlq_lambda0_equals_reference_ivfpq(cuda, ref, d, C, E, M, P):
    from vector_line_quantization_b200 import data, index as vi, ops

    k = 10
    xt = data.sift_like(20000, d=d, kc=256, seed=61)
    xb = data.sift_like(30000, d=d, kc=256, seed=62)
    xq = data.sift_like(100, d=d, kc=256, seed=63)
    ivf = ref.RefIVFPQ(d, C, M)           # the reference's own train: Clustering + ProductQuantizer on the CPU
    ivf.train(xt)
    cent, pq = ivf.codebooks()
    ivf.add(xb)
    Dr, Ir = ivf.search(xq, k, P)

    import torch

    edge, ed2 = ops.knn_graph(torch.from_numpy(cent).to(cuda), E)
    res = vi.StandardGpuResources(0)
    idx = vi.GpuIndexIVFPQ(res, d, C, M, 8, E, 1)
    idx.setCodebooks(cent, edge.cpu().numpy(), ed2.cpu().numpy(), np.zeros(1, np.float32), pq)
    idx.add(xb)
    # encode parity: the union of the E line lists of centroid c holds exactly the reference's list c, same PQ codes
    for c in range(C):
        ids_r, codes_r = ivf.get_list(c)
        got = {}
        for e in range(E):
            codes, las, ids = idx.getList(c * E + e)
            assert np.all(las == 0)
            for i, code in zip(ids, codes):
                got[int(i)] = code
        common = [i for i in ids_r if int(i) in got]
        assert len(common) >= len(ids_r) - 2 and len(got) <= len(ids_r) + 2  # coarse near-ties may move a vector
        same = np.mean([np.array_equal(got[int(i)], codes_r[j]) for j, i in enumerate(ids_r) if int(i) in got])
        assert same > 0.995
    # search parity
    idx.setNumProbes(P)
    idx.w1_ = P * E
    idx.setListCap(1 << 20)
    D, I = idx.search(xq, k)
    qn = np.sum(xq.astype(np.float64) ** 2, axis=1, keepdims=True)
    assert (I == Ir).mean() > 0.97
    same = I == Ir
    np.testing.assert_allclose((D + qn)[same], Dr[same], rtol=2e-4)
    overlap = np.mean([len(set(a) & set(b)) / k for a, b in zip(I, Ir)])
    assert overlap > 0.99
