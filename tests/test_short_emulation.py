"""CPU: lane-by-lane emulation of b200::csr_short_kernel (cudalibrarysamples_b200/csrc/spmv_csr_short.cu): a warp per 32
consecutive rows, passes of 32 * STEPS non-zeros starting at a multiple of 32, products parked at slot i + (i >> 5), lane l
adds row l's slice of every pass.  Pins the index arithmetic (range ends of the padding lanes, pass borders inside a row,
rows longer than one pass, empty rows) against the oracle without a GPU; the kernel itself runs under test_parity_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as O

STEPS = 8
CAP = 32 * STEPS


def slot(i):
    return i + (i >> 5)


def emulate(off, col, val, x, y0, alpha, beta):
    off = off.astype(np.int64)
    rows = off.size - 1
    y = y0.astype(np.float64).copy()
    lane = np.arange(32)
    for blk in range((rows + 31) // 32):
        row = blk * 32 + lane
        live = row < rows
        rb = off[np.minimum(row, rows)]
        re = np.where(live, off[np.minimum(row + 1, rows)], rb)
        b, e = int(rb[0]), int(re[31])
        s = np.zeros(32)
        p0 = b & ~31
        while p0 < e:
            sp = np.full(CAP + STEPS, np.nan)
            for k in range(STEPS):
                if p0 + k * 32 < e:
                    i = p0 + k * 32 + lane
                    inside = i < e
                    ii = np.minimum(i, e - 1)
                    sp[slot(k * 32 + lane)] = np.where(inside, val[ii] * x[col[ii]], 0.0)
            lo, hi = np.maximum(rb, p0) - p0, np.minimum(re, p0 + CAP) - p0
            for l in range(32):
                for i in range(int(lo[l]), int(hi[l])):
                    assert not np.isnan(sp[slot(i)])
                    s[l] += sp[slot(i)]
            p0 += CAP
        for l in range(32):
            if live[l]:
                r = int(row[l])
                y[r] = alpha * s[l] if beta == 0 else alpha * s[l] + beta * y[r]
    return y


def build(lens, cols, seed):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    parts = [np.sort(rng.choice(cols, size=int(l), replace=False)) for l in lens if l > 0]
    col = (np.concatenate(parts) if parts else np.zeros(0, int)).astype(np.int32)
    return off, col, rng.uniform(-1, 1, int(off[-1]))


PROFILES = {
    "stencil_like": [5] * 100 + [3, 4] * 10,
    "seven_point": [7] * 77,
    "rows_of_8": [8] * 64,
    "rows_of_32": [32] * 40,
    "mixed_short": [1, 2, 3, 0, 5, 1, 1, 0, 0, 7, 16, 31, 32] * 11,
    "all_empty": [0] * 70,
    "one_row": [3],
    "leading_and_trailing_empty": [0] * 40 + [5, 0, 30, 0, 0, 2] + [0] * 50,
    "longer_than_a_pass": [700, 3, 0, 300, 255, 257, 256, 1, 1000],          # forced mode: the multi-pass path
    "pass_border_inside_rows": [100] * 30,
}


@pytest.mark.parametrize("name", list(PROFILES))
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-2.0, 0.5)])
def test_short_kernel_bookkeeping(name, alpha, beta):
    lens = np.array(PROFILES[name])
    off, col, val = build(lens, 5000, 11)
    x, y0 = O.uniform(1, 5000), O.uniform(2, lens.size)
    want = O.spmv_csr(off, col, val, x, y0, alpha, beta)
    got = emulate(off, col, val, x, y0, alpha, beta)
    assert np.linalg.norm(got - want) <= 1e-13 * max(np.linalg.norm(want), 1e-300)


def test_short_kernel_on_the_cg_sample_operator():
    off, col, val = O.gen_stencil5(37)                       # cg_example.c:71-128
    n = 37 * 37
    x, y0 = O.uniform(3, n), O.uniform(4, n)
    assert np.linalg.norm(emulate(off, col, val, x, y0, 0.75, -1.0) - O.spmv_csr(off, col, val, x, y0, 0.75, -1.0)) < 1e-12
