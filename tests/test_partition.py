"""CPU: the restated tile partition (oracle/partition_ref.py) keeps its invariants on awkward row-length profiles."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.partition_ref import check_partition, csr_partition

TILE, LONG = 2048, 512


def offsets(lens):
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


CASES = {
    "uniform16": np.full(5000, 16),
    "empty": np.zeros(10000, int),
    "single_huge": np.array([100000]),
    "huge_then_tiny": np.concatenate([[50000], np.ones(3000, int), [0] * 500, [700], [511], [512], [513]]),
    "all_just_below_long": np.full(300, LONG - 1),
    "all_exactly_long": np.full(300, LONG),
    "tile_sized": np.full(40, TILE),
    "tile_minus_one": np.full(40, TILE - 1),
    "alternating": np.tile([0, 1, 4095, 0, 0, 3], 200),
    "one_row_one_nnz": np.array([1]),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("base", [0, 1])
def test_invariants(name, base):
    off = offsets(CASES[name]) + base
    t = csr_partition(off, base, TILE, LONG)
    assert check_partition(t, off, base, TILE, LONG)


def test_short_rows_are_never_cut():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, LONG, 20000)
    off = offsets(lens)
    t = csr_partition(off, 0, TILE, LONG)
    assert np.array_equal(t[:, 1], off[t[:, 0]])


def test_split_row_arrival_counts_are_consistent():
    """Every cut row R must be covered by exactly b2-b1+1 tiles with b1 = g(R)//TILE, b2 = (g(R)+len)//TILE --
    the numbers split_row_arrive() in spmv_csr.cu derives in-kernel."""
    lens = CASES["huge_then_tiny"]
    off = offsets(lens).astype(np.int64)
    t = csr_partition(off, 0, TILE, LONG).astype(np.int64)
    rows = lens.size
    contrib = {}
    for b in range(t.shape[0] - 1):
        (rs, ns), (re, ne) = t[b], t[b + 1]
        if rs < rows and ns > off[rs]:
            contrib.setdefault(int(rs), []).append(("head", b))
        r_first = rs + (1 if (rs < rows and ns > off[rs]) else 0)
        if re < rows and re >= r_first and ne > off[re]:
            contrib.setdefault(int(re), []).append(("tail", b))
    assert contrib, "the case must contain split rows"
    for R, lst in contrib.items():
        g0 = R + off[R]
        g1 = R + off[R + 1]
        b1, b2 = g0 // TILE, g1 // TILE
        assert len(lst) == b2 - b1 + 1
        assert lst[0] == ("tail", b1)
        assert [b for _, b in lst[1:]] == list(range(b1 + 1, b2 + 1))


def test_rmat_partition():
    off, col, val = O.rmat_csr(50000, avg_nnz=16, seed=11, val_seed=12)
    t = csr_partition(off, 0, TILE, LONG)
    assert check_partition(t, off, 0, TILE, LONG)
