"""The C-ABI library must load and export every symbol include/*.h declares (no compute calls: runs without a GPU)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(r"\b(vlq[a-z0-9_]*)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    from vector_line_quantization_b200 import build

    return build.build_all()


def test_cuda_library_exports_header(built):
    lib = ctypes.CDLL(built[0])
    syms = declared_symbols(os.path.join(ROOT, "include", "vlq_b200.h"))
    assert len(syms) > 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from vector_line_quantization_b200 import _abi

    syms = declared_symbols(os.path.join(ROOT, "include", "vlq_b200.h"))
    assert sorted(_abi.SIGNATURES) == syms
    handle = _abi.lib()
    assert handle.vlq_version().startswith(b"vlq_b200")
    assert handle.vlq_error_string(-1) == b"vlq: invalid argument"


def test_host_library_exports_header(built):
    hdr = os.path.join(ROOT, "include", "vlq_index_c.h")
    if built[1] is None or not os.path.exists(hdr):
        pytest.skip("host layer not built yet")
    lib = ctypes.CDLL(built[1])
    missing = [s for s in declared_symbols(hdr) if not hasattr(lib, s)]
    assert not missing, missing


def test_no_oracle_in_product_path():
    """the product tree never references oracle/ (a CPU fallback would void every parity claim)"""
    pkg = os.path.join(ROOT, "vector_line_quantization_b200")
    offenders = []
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isdir(path) or path.endswith((".so", ".o", ".pyc", ".log")):
            continue
        txt = open(path, errors="ignore").read()
        if re.search(r"pyoracle|vlq_oracle|libfaiss_ref|from oracle|import oracle", txt):
            offenders.append(path)
    assert not offenders, offenders


def test_invalid_arguments_return_codes():
    from vector_line_quantization_b200 import _abi

    h = _abi.lib()
    assert h.vlq_row_norms(None, 4, 8, None, None) == -1
    assert h.vlq_l2_assign(None, 1, 8, None, None, 4, 1, None, None, None) == -1
    assert h.vlq_merge_topk(None, None, 2, 4, 8, None, None, None) == -1
    assert h.vlq_select_rows(None, 1, 10, 10, 2000, None, None, None, None) == -1


def test_coarse_route_dispatch_rule(built):
    """host-side dispatch of the coarse stage (no GPU): the matrix-free route is supported for d % 4 == 0, d <= 128,
    max(num_buckets, 32 P, P E) <= 4096, W <= 1024 and preferred up to 256 KiB of centroid rows per query"""
    lib = ctypes.CDLL(built[0])
    f = lambda name, *a: getattr(lib, name)(*[ctypes.c_int(x) for x in a])  # noqa: E731
    assert lib.vlq_tc_num_buckets(ctypes.c_int(65536)) == 2048 and lib.vlq_tc_num_buckets(ctypes.c_int(130)) == 8
    # (d, C, P, E, W)
    assert f("vlq_coarse_exact_supported", 128, 65536, 64, 32, 256) == 1
    assert f("vlq_coarse_exact_preferred", 128, 65536, 64, 32, 256) == 0   # 2 MiB of rows per query: matrix route
    assert f("vlq_coarse_exact_preferred", 128, 65536, 8, 32, 32) == 1     # 256 KiB: matrix-free
    assert f("vlq_coarse_exact_preferred", 128, 65536, 16, 32, 64) == 0
    assert f("vlq_coarse_exact_preferred", 96, 65536, 8, 64, 32) == 0      # P (32 + E) d 4 = 288 KiB
    assert f("vlq_coarse_exact_preferred", 96, 65536, 4, 64, 16) == 1      # 144 KiB
    assert f("vlq_coarse_exact_supported", 130, 65536, 8, 32, 32) == 0     # d % 4
    assert f("vlq_coarse_exact_supported", 256, 65536, 8, 32, 32) == 0     # d > 128
    assert f("vlq_coarse_exact_supported", 128, 65536, 200, 32, 256) == 0  # 32 P > 4096
    assert f("vlq_coarse_exact_supported", 128, 1 << 20, 8, 32, 32) == 0   # 32768 buckets
    assert f("vlq_coarse_exact_supported", 128, 65536, 0, 32, 32) == 0


def test_query_tile_split():
    """equal query tiles of at most 5120 rows, multiples of 256 (the split GpuIndexIVFPQ::search makes per page)"""
    from vector_line_quantization_b200.ops import _tile_rows

    assert [_tile_rows(n, 5120) for n in (1, 100, 5120, 5121, 10000, 32768)] == [1, 100, 5120, 2816, 5120, 4864]
    for n in (5121, 10000, 12345, 32768, 40000):
        t = _tile_rows(n, 5120)
        assert t <= 5120 and t % 256 == 0 and -(-n // t) == -(-n // 5120)  # no more tiles than the cap requires
