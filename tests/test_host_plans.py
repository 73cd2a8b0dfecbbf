"""Host-side launch computations of the library, checked without a GPU through the C ABI:
the row tiling of the tensor-core coarsest level (a wrong plan would silently skip or repeat query rows) and the
multiply-high / shift division the gather kernels use for their index arithmetic."""
import ctypes as C
import random

import pytest

from casmtr_b200 import _lib


def _plan(rows, bh, n_sm):
    out = (C.c_int * 3)()
    assert _lib.lib().casmtr_plan_dense_tiles(rows, bh, n_sm, out) == 0
    return out[0], out[1], out[2]


@pytest.mark.parametrize('n_sm', [148, 132, 1, 8])
def test_dense_tiles_cover_every_row_exactly_once(n_sm):
    rnd = random.Random(7)
    cases = [(676, 16), (676, 32), (676, 64), (676, 128), (300, 64), (64, 8), (65, 8), (1, 1), (704, 16), (128, 148), (4096, 3)]
    cases += [(rnd.randrange(1, 3000), rnd.randrange(1, 400)) for _ in range(300)]
    for rows, bh in cases:
        n_big, n_small, rows_small = _plan(rows, bh, n_sm)
        assert n_big >= 0 and n_small >= 0 and n_big * 64 <= rows
        left = rows - 64 * n_big
        if left > 0:                                                  # (with nothing left, trailing tiles own no rows: the kernel exits on n_rows <= 0)
            assert n_small >= 1 and 1 <= rows_small <= 64 and rows_small % 2 == 0, (rows, bh, n_big, n_small, rows_small)
            assert n_small * rows_small >= left, (rows, bh, n_big, n_small, rows_small)
        # the kernel's own mapping: tile t owns [row0, row0 + n_rows)
        covered = 0
        for t in range(n_big + n_small):
            row0 = t * 64 if t < n_big else n_big * 64 + (t - n_big) * rows_small
            n_rows = min(64 if t < n_big else rows_small, rows - row0)
            if n_rows <= 0:
                continue
            assert row0 == covered, (rows, bh, t)
            covered += n_rows
        assert covered == rows, (rows, bh, n_big, n_small, rows_small)


def test_dense_tiles_832_one_pair_is_one_long_and_one_short_wave():
    assert _plan(676, 16, 148) == (9, 9, 12)            # 144 CTAs of 64 rows, then 144 of 12
    n_big, n_small, rows_small = _plan(676, 64, 148)    # 4 pairs: plain tiling (4.8 waves either way)
    assert (n_big, n_small, rows_small) == (10, 1, 36)


def test_fastdiv_is_exact():
    rnd = random.Random(3)
    divisors = list(range(1, 130)) + [169, 208, 676, 2704, 10816, 43264, 173056, 1 << 20, (1 << 31) - 1] + [rnd.randrange(1, 1 << 31) for _ in range(200)]
    for d in divisors:
        out = (C.c_uint * 2)()
        assert _lib.lib().casmtr_fastdiv(d, out) == 0
        mul, shr = out[0], out[1]
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - 2] + [rnd.randrange(1 << 31) for _ in range(200)]:
            if n < 0 or n >= 1 << 31:
                continue
            q = ((n * mul) >> 32) >> shr if mul else n
            assert q == n // d, (d, n, q)


def test_plan_rejects_bad_arguments():
    out = (C.c_int * 3)()
    assert _lib.lib().casmtr_plan_dense_tiles(0, 1, 148, out) != 0
    assert _lib.lib().casmtr_fastdiv(0, (C.c_uint * 2)()) != 0
