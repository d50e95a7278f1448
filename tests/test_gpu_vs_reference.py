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


@pytest.mark.gpu
@pytest.mark.parametrize("nedge", [1, 4])
def test_copy_from_reference_cpu_index_and_back(cuda, oracle, nedge):
    """f4, no oracle in between: the UNMODIFIED reference CPU IndexIVFPQ (oracle/_ref) is trained and filled on the CPU,
    its codebooks and inverted lists go through GpuIndexIVFPQ::copyFrom, and GpuIndexIVFPQ::search must return what
    IndexIVFPQ::search returns (gpu/GpuIndexIVFPQ.cu:169-232); copyTo gives the same lists back (:234-281)"""
    if not oracle.ref_available():
        pytest.skip("reference library not built")
    from vector_line_quantization_b200 import data, index as vi

    d, nlist, M, k, nprobe = 64, 64, 8, 10, 8
    xt = data.sift_like(20000, d=d, kc=256, seed=1)
    xb = data.sift_like(30000, d=d, kc=256, seed=2)
    xq = data.sift_like(200, d=d, kc=256, seed=3)
    ref = oracle.RefIVFPQ(d, nlist, M, 8)
    ref.train(xt)
    ref.add(xb)
    Dr, Ir = ref.search(xq, k, nprobe)
    coarse, pq = ref.codebooks()
    cpu = vi.CpuIndexIVFPQ(d, nlist, M, 8)
    cpu.set_codebooks(coarse, pq)
    for l in range(nlist):
        cpu.set_list(l, *ref.get_list(l))
    assert cpu.ntotal == len(xb)
    res = vi.StandardGpuResources(0)
    gpu = vi.GpuIndexIVFPQ(res, d, nlist, M, 8, nedge, 16)
    gpu.copyFrom(cpu)
    assert gpu.ntotal == len(xb)
    gpu.setNumProbes(nprobe)
    gpu.w1_ = nprobe * nedge  # every line of the probed centroids: exactly IndexIVFPQ::search(nprobe)
    gpu.setListCap(1 << 20)  # the reference's CPU search has no 1024-entry cap per list
    D, I = gpu.search(xq, k)
    qn = (xq.astype(np.float64) ** 2).sum(1, keepdims=True)  # the GPU path omits ||q||^2
    same = I == Ir
    assert same.mean() > 0.97  # (a coarse near-tie at the nprobe boundary changes which lists a query scans)
    np.testing.assert_allclose((D + qn)[same], Dr[same], rtol=2e-4)
    assert np.mean([len(set(a) & set(b)) / k for a, b in zip(I, Ir)]) > 0.99
    back = vi.CpuIndexIVFPQ(d, nlist, M, 8)
    gpu.copyTo(back)
    c2, p2 = back.codebooks()
    assert np.array_equal(c2, coarse) and np.array_equal(p2, pq) and back.ntotal == len(xb)
    for l in range(nlist):
        i0, k0 = ref.get_list(l)
        i1, k1 = back.get_list(l)
        assert np.array_equal(i0, i1) and np.array_equal(k0, k1)
