"""Address arithmetic of the bank-skewed scan tables (csrc/search.cu, scan_async_kernel<M, true>), restated in numpy:
the invariants the kernel relies on, checked exhaustively over lanes, steps and codes.  CPU only.

Row `code` of the table is 64 words; word c (c < 31 + M) holds T3[c % M][code].  Lane l has column base
cb = M*(l // M) + l % M and, at step s, looks up sub-quantizer (s + l) % M of its entry at word cb + s of row code;
its code bytes were rotated left by l % M so that byte s is the code of that sub-quantizer."""
import numpy as np
import pytest

ROW_WORDS = 64


@pytest.mark.parametrize("M", [8, 16])
def test_skewed_lookup_is_conflict_free_and_addresses_the_right_table(M):
    lanes = np.arange(32)
    lp, grp = lanes % M, lanes // M
    cb = M * grp + lp
    assert (4 * cb).max() < 256  # the lane offset must fit the low byte of `code << 8 | lofs`
    rng = np.random.RandomState(M)
    codes = rng.randint(0, 256, size=(32, M))  # one entry per lane
    for s in range(M):
        col = cb + s
        assert col.max() < 31 + M <= ROW_WORDS  # inside the filled part of the row
        sub = (s + lp) % M
        assert np.array_equal(col % M, sub)  # word col of any row belongs to the sub-quantizer the lane works on
        rotated = np.stack([np.roll(codes[l], -int(lp[l])) for l in range(32)])  # byte s after the lane's rotation
        assert np.array_equal(rotated[:, s], codes[lanes, sub])
        word = rotated[:, s] * ROW_WORDS + col  # word address inside the table
        banks = word % 32
        assert len(set(banks.tolist())) == 32  # 32 lanes, 32 different banks, whatever the codes are
    # every lane visits every sub-quantizer exactly once
    for l in range(32):
        assert sorted(((np.arange(M) + lp[l]) % M).tolist()) == list(range(M))


def test_table_fill_pattern():
    """the fill copies whole float4s: float4 q of row `code` = source float4 (q mod M/4) of the code-major T3 row"""
    for M in (8, 16):
        q4 = (31 + M + 3) // 4
        words = np.arange(4 * q4)
        src = 4 * ((words // 4) % (M // 4)) + words % 4
        assert np.array_equal(src, words % M)
        assert 4 * q4 <= ROW_WORDS and 4 * q4 >= 31 + M
