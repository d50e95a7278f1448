"""The built sm_100a objects really contain the Blackwell instructions DESIGN.md claims (B200_PROFILING.md, "What proves
a Blackwell-native kernel"): tcgen05 MMA + TMEM loads + bulk TMA copies + mbarriers in the assignment GEMM, the TMA bulk
prefetch in the long-list scan, and no legacy HMMA anywhere.  Needs nvcc / cuobjdump only (no GPU)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "vector_line_quantization_b200", "build", "obj")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    from vector_line_quantization_b200 import build

    build.build_cuda()
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

    def dump(name):
        return subprocess.run([tool, "-sass", os.path.join(OBJ, name)], capture_output=True, text=True,
                              stdin=subprocess.DEVNULL, timeout=300).stdout

    return dump


def test_assignment_gemm_is_tcgen05_with_tma(sass):
    s = sass("assign_tc.o")
    assert "sm_100a" in s
    for pat, what in ((r"\bUTC[A-Z]*MMA\b", "tcgen05.mma"), (r"\bLDTM\b", "tcgen05.ld"), (r"\bUBLKCP\b", "cp.async.bulk"),
                      (r"\bSYNCS\b", "mbarrier")):
        assert re.search(pat, s), "no %s in assign_tc.o" % what
    assert not re.search(r"\bHMMA\b", s), "legacy mma.sync path in assign_tc.o"


def test_long_list_scan_uses_tma_bulk_prefetch(sass):
    s = sass("scan_long.o")
    assert re.search(r"\bUBLKPF\b", s), "no cp.async.bulk.prefetch.L2 in scan_long.o"
    assert len(re.findall(r"\bLDG\.E(\.[A-Z0-9]+)*\.128\b", s)) > 0  # 16-byte vectorised code loads


def test_no_legacy_tensor_path_anywhere(sass):
    for name in sorted(os.listdir(OBJ)):
        if name.endswith(".o"):
            assert not re.search(r"\b(HMMA|HGMMA)\b", sass(name)), name
