"""CPU: lane-level numpy emulation of csr_generic_kernel / sell_generic_kernel (csrc/spmv_generic.cu) against the oracle.

The emulation follows the kernel statement by statement -- the warp-uniform outer loop over groups of 32 / L rows, the lanes
striding through their row, the shuffle tree with `width = L` (a lane whose partner would fall outside its L-lane segment
gets its own value back, as __shfl_down_sync does), lane 0 of a segment writing y -- so it pins the index arithmetic and
the reduction pattern: every row written exactly once, with the complete row sum, for every lane count and for grids
smaller than the matrix (several trips of the loop).  The GPU parity proper is tests/test_generic_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as O


def shfl_down(v, o, width):
    lane = np.arange(32)
    src = np.where((lane % width) + o < width, lane + o, lane)
    return v[src]


def emulate_csr_generic(off, col, val, x, y, alpha, beta, base, lanes_log2, ctas, block=256):
    lanes = 1 << lanes_log2
    rows = off.size - 1
    rows_per_warp = 32 >> lanes_log2
    nwarps = ctas * block // 32
    written = np.zeros(rows, int)
    lane = np.arange(32)
    sub = lane & (lanes - 1)
    for warp in range(nwarps):
        w0 = warp * rows_per_warp
        while w0 < rows:                                   # warp-uniform trip count
            row = w0 + (lane >> lanes_log2)
            s = np.zeros(32)
            for l in range(32):
                if row[l] < rows:
                    beg, end = int(off[row[l]]) - base, int(off[row[l] + 1]) - base
                    k = beg + sub[l]
                    while k < end:
                        s[l] += val[k] * x[int(col[k]) - base]
                        k += lanes
            o = lanes >> 1
            while o > 0:
                s = s + shfl_down(s, o, lanes)
                o >>= 1
            for l in range(32):
                if row[l] < rows and sub[l] == 0:
                    r = row[l]
                    y[r] = alpha * s[l] if beta == 0 else alpha * s[l] + beta * y[r]
                    written[r] += 1
            w0 += nwarps * rows_per_warp
    return written


@pytest.mark.parametrize("lanes_log2", [2, 3, 4, 5])
@pytest.mark.parametrize("base", [0, 1])
def test_csr_generic_lane_logic(lanes_log2, base):
    rows = 700
    off, col, val = O.rmat_csr(rows, avg_nnz=9, seed=11, val_seed=12)
    x, y0 = O.uniform(13, rows), O.uniform(14, rows)
    want = O.spmv_csr(off, col, val, x, y0, -1.5, 0.5)
    y = y0.copy()
    # one CTA = 8 warps: far fewer rows per sweep than the matrix has -> several trips of the outer loop
    written = emulate_csr_generic(off.astype(np.int64) + base, col.astype(np.int64) + base, val, x, y, -1.5, 0.5, base, lanes_log2, ctas=1)
    assert np.all(written == 1)
    assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)


def test_csr_generic_rows_not_a_multiple_of_the_group_and_empty_rows():
    lens = np.array([0, 5, 0, 0, 33, 1, 64, 0, 7, 129, 0])            # 11 rows: the last group of every lane count is partial
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rng = np.random.default_rng(3)
    col = rng.integers(0, 50, off[-1]).astype(np.int64)
    val = rng.uniform(-1, 1, off[-1])
    x, y0 = rng.uniform(-1, 1, 50), rng.uniform(-1, 1, lens.size)
    want = O.spmv_csr(off.astype(np.int32), col.astype(np.int32), val, x, y0, 2.0, 0.0)
    for lanes_log2 in (2, 3, 4, 5):
        y = np.full(lens.size, np.nan)                                 # beta = 0 never reads y
        written = emulate_csr_generic(off, col, val, x, y, 2.0, 0.0, 0, lanes_log2, ctas=1)
        assert np.all(written == 1)
        assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)


def emulate_sell_generic(rows, S, so, sc, sv, x, y, alpha, beta, base):
    for row in range(rows):
        s, r = divmod(row, S)
        beg, end = int(so[s]) - base, int(so[s + 1]) - base
        width = (end - beg) // S
        acc = 0.0
        for k in range(width):
            i = beg + k * S + r
            c = int(sc[i]) - base
            if c < 0:
                continue
            acc += sv[i] * x[c]
        y[row] = alpha * acc if beta == 0 else alpha * acc + beta * y[row]


@pytest.mark.parametrize("S", [2, 7, 32])
def test_sell_generic_index_arithmetic(S):
    rows = 203                                                         # partial last slice
    off, col, val = O.rmat_csr(rows, avg_nnz=5, seed=21, val_seed=22)
    so, sc, sv = O.csr_to_sell(off, col, val, S)
    x, y0 = O.uniform(23, rows), O.uniform(24, rows)
    y = y0.copy()
    emulate_sell_generic(rows, S, so.astype(np.int64), sc.astype(np.int64), sv, x, y, 1.0, 2.0, 0)
    want = O.spmv_csr(off, col, val, x, y0, 1.0, 2.0)
    assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)
    assert np.linalg.norm(O.spmv_sell(rows, S, so, sc, sv, x, y0, 1.0, 2.0) - want) <= 1e-13 * np.linalg.norm(want)
