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
