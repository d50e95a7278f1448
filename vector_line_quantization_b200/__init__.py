"""vlq-b200: B200-native (sm_100a) implementation of the vector-line-quantization hot path.

Layout
  csrc/    hand-written CUDA kernels + the C-ABI (include/vlq_b200.h)      -> lib/libvlq_b200.so
  host/    C++ host layer mirroring faiss::Index / GpuIndexFlatL2 / GpuIndexIVFPQ(VLQ) / IndexProxy,
           plus a C wrapper (include/vlq_index_c.h)                        -> lib/libvlq_host.so
  _abi.py  ctypes binding of the C-ABI          ops.py    device-resident operator layer (torch tensors as plumbing)
  index.py Python mirror of the host classes    data.py   deterministic synthetic SIFT/DEEP-shaped generators

There is no CPU fallback anywhere in this package: importing the ops without the built CUDA library raises.
"""
__version__ = "0.1.0"
