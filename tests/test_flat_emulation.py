"""CPU: lane-by-lane emulation of b200::csr_flat_kernel + csr_flat_fixup_kernel (cudalibrarysamples_b200/csrc/
spmv_csr_flat.cu) on the plan restated in oracle/partition_ref.py::flat_plan.

Pins the ALGORITHM without a GPU: the end-lane mask per step, first-row butterfly + segmented scan, the row lookup through
chunk_run / nzrow, the empty-row pass, the deferred first row end of every warp chunk, the per-CTA stitch and the
cross-CTA fix-up -- every row must be written exactly once.  The kernel itself runs under tests/test_parity_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.partition_ref import flat_plan

STEPS, CHUNK = 8, 256
WARPS = 4                      # B200_FLAT_WARPS: warp chunks stitched per CTA (the plan itself does not depend on it)
CTA = CHUNK * WARPS
CTA_WORDS = CTA // 32


def axpby(alpha, s, beta, yv):
    return alpha * s if beta == 0 else alpha * s + beta * yv


def emulate(off, col, val, x, y0, alpha, beta, warps=None):
    global WARPS, CTA, CTA_WORDS
    if warps is not None:
        WARPS, CTA, CTA_WORDS = warps, CHUNK * warps, CHUNK * warps // 32
    off = off.astype(np.int64)
    rows, nnz = off.size - 1, int(off[-1])
    y = y0.astype(np.float64).copy()
    written = np.zeros(rows, int)

    def store(r, v):
        assert 0 <= r < rows
        written[r] += 1
        y[r] = axpby(alpha, v, beta, y[r])

    if nnz == 0:
        for r in range(rows):
            store(r, 0.0)
        return y, written
    mask, chunk_run, nzrow, (nruns, _, _) = flat_plan(off, 0)
    prod = val * x[col]
    nctas = (nnz + CTA - 1) // CTA
    cta_first, cta_last, cta_flags = np.full(nctas, np.nan), np.full(nctas, np.nan), np.zeros(nctas, int)
    lane = np.arange(32)
    for cta in range(nctas):
        sFirst, sLast, sFrow = np.zeros(WARPS), np.zeros(WARPS), np.full(WARPS, -1)
        for warp in range(WARPS):
            c = cta * WARPS + warp
            n0 = c * CHUNK
            acc, first, last, frow = np.zeros(32), 0.0, 0.0, -1
            if n0 < nnz:
                n1 = min(n0 + CHUNK, nnz)
                mreg = mask[c * STEPS:(c + 1) * STEPS]
                run = int(chunk_run[c])
                for k in range(STEPS):
                    if n0 + k * 32 >= n1:
                        break
                    e = n0 + k * 32 + lane
                    pk = np.where(e < n1, prod[np.minimum(e, nnz - 1)], 0.0)
                    m = int(mreg[k])
                    if m == 0:
                        acc = acc + pk
                        continue
                    e1, ek = (m & -m).bit_length() - 1, m.bit_length() - 1
                    t1 = float((acc + np.where(lane <= e1, pk, 0.0)).sum())
                    q = np.where((lane > e1) & (lane <= ek), pk, 0.0)
                    if m & (m - 1):
                        below = np.array([m & ((1 << int(l)) - 1) for l in lane])
                        dist = np.where((lane > e1) & (lane <= ek), lane - np.array([int(v).bit_length() for v in below]), 0)
                        d = 1
                        while d < 32 and np.any(dist >= d):
                            t = np.concatenate([q[:d], q[:-d]])
                            q = np.where(dist >= d, q + t, q)
                            d <<= 1
                    res = np.where(lane == e1, t1, q)
                    is_end = ((m >> lane) & 1).astype(bool)
                    j = run + np.array([bin(m & ((1 << int(l)) - 1)).count("1") for l in lane])
                    row = np.where(is_end, nzrow[np.minimum(j + 1, nruns + 1)], 0)
                    deferred = -1
                    if frow < 0:
                        first, frow, deferred = float(res[e1]), int(row[e1]), e1
                    for l in lane[is_end]:
                        if l != deferred:
                            store(int(row[l]), float(res[l]))
                    run += bin(m).count("1")
                    acc = np.where(lane > ek, pk, 0.0)
                el = n1 - 1 - n0
                if not (int(mreg[el >> 5]) >> (el & 31)) & 1:
                    last = float(acc.sum())
                else:
                    assert not np.any(acc != 0)
            if frow >= 0:
                sFirst[warp], sLast[warp] = first, last
            else:
                sFirst[warp], sLast[warp] = last, 0.0
            sFrow[warp] = frow
        starts_row = cta == 0 or (int(mask[cta * CTA_WORDS - 1]) >> 31) & 1
        running, has = 0.0, False
        for w in range(WARPS):
            if sFrow[w] >= 0:
                tot = running + sFirst[w]
                if not has and not starts_row:
                    cta_first[cta] = tot
                else:
                    store(int(sFrow[w]), tot)
                running, has = sLast[w], True
            else:
                running += sFirst[w]
        if not has:
            cta_first[cta], cta_last[cta] = running, 0.0
        else:
            cta_last[cta] = running
        cta_flags[cta] = int(has)
    for r in range(rows):                                        # csr_flat_fixup_kernel, part 2: the empty rows
        if off[r] == off[r + 1]:
            store(r, 0.0)
    for t in range(nctas - 1):                                   # csr_flat_fixup_kernel, part 1: rows crossing CTA borders
        if (int(mask[(t + 1) * CTA_WORDS - 1]) >> 31) & 1:
            continue
        has = cta_flags[t] != 0
        starts_row = t == 0 or (int(mask[t * CTA_WORDS - 1]) >> 31) & 1
        if not has and not starts_row:
            continue
        s = cta_last[t] if has else cta_first[t]
        u = t + 1
        while cta_flags[u] == 0:
            s += cta_first[u]
            u += 1
        s += cta_first[u]
        store(int(nzrow[chunk_run[u * WARPS] + 1]), s)
    return y, written


def build(lens, cols, seed):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    parts = [np.sort(rng.choice(cols, size=int(l), replace=False)) for l in lens if l > 0]
    col = (np.concatenate(parts) if parts else np.zeros(0, int)).astype(np.int32)
    val = rng.uniform(-1, 1, int(off[-1]))
    return off, col, val


PROFILES = {
    "long_rows_mixed": [700, 64, 100, 3000, 0, 0, 65, 2048, 511, 513, 1200, 90, 64, 64, 5000, 130],
    "single_huge_row": [20000],
    "huge_with_empties_around": [0, 0, 9000, 0, 0, 0, 4100, 0],
    "rows_of_exactly_32": [32] * 200,
    "rows_of_exactly_256": [256] * 20,            # row ends coincide with warp-chunk borders
    "rows_of_exactly_2048": [2048] * 4,           # row ends coincide with CTA borders
    "dense_then_sparse": [3000, 2500, 800] + [3] * 400 + [1500, 0, 0, 900],
    "long_row_then_40_empty_rows": [2500] + [0] * 40 + [1500] + [0] * 70 + [30, 2000],
    "chunk_border_plus_minus_one": [255, 257, 256, 1, 255, 512, 513, 31, 33, 32, 992],
    "rmat_like_block": [500, 158, 158, 50, 158, 50, 50, 16, 158, 50, 50, 16, 50, 16, 16, 5] * 6,
    "short_rows_only": [1, 2, 3, 0, 5, 1, 1, 0, 0, 7] * 90,
    "leading_and_trailing_empty": [0] * 50 + [5, 0, 300, 0, 0, 2] + [0] * 70,
    "all_empty": [0] * 100,
    "one_nonzero": [0, 0, 1, 0],
    "cta_spanning_rows": [5000, 1, 4095, 2049, 2047, 6144, 3],
}


@pytest.mark.parametrize("name", list(PROFILES))
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-2.0, 0.5)])
def test_flat_kernel_bookkeeping(name, alpha, beta):
    lens = np.array(PROFILES[name])
    cols = 30000
    off, col, val = build(lens, cols, 7)
    x, y0 = O.uniform(1, cols), O.uniform(2, lens.size)
    want = O.spmv_csr(off, col, val, x, y0, alpha, beta)
    got, written = emulate(off, col, val, x, y0, alpha, beta)
    assert np.all(written == 1), "every row is written exactly once"
    assert np.linalg.norm(got - want) <= 1e-12 * max(np.linalg.norm(want), 1e-300)


@pytest.mark.parametrize("seed", range(6))
def test_flat_kernel_random_structures(seed):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(50, 400))
    kind = rng.integers(0, 5, n)
    lens = np.where(kind == 0, 0, np.where(kind == 1, rng.integers(1, 6, n), np.where(kind == 2, rng.integers(6, 80, n),
                    np.where(kind == 3, rng.integers(80, 700, n), rng.integers(700, 6000, n)))))
    if seed % 2:
        lens[rng.integers(0, n, n // 3)] = 0
    cols = 50000
    off, col, val = build(lens, cols, seed)
    x, y0 = O.uniform(11 + seed, cols), O.uniform(12 + seed, n)
    want = O.spmv_csr(off, col, val, x, y0, 0.75, 1.25)
    got, written = emulate(off, col, val, x, y0, 0.75, 1.25)
    assert np.all(written == 1)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


@pytest.mark.parametrize("warps", [8, 4, 2])
def test_flat_kernel_on_rmat(warps):
    off, col, val = O.rmat_csr(40000, avg_nnz=16, seed=3, val_seed=4)
    x, y0 = O.uniform(5, 40000), O.uniform(6, 40000)
    want = O.spmv_csr(off, col, val, x, y0, 1.5, -0.25)
    got, written = emulate(off, col, val, x, y0, 1.5, -0.25, warps=warps)
    assert np.all(written == 1)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


def test_flat_plan_invariants():
    off = O.rmat_csr(30000, avg_nnz=16, seed=5, val_seed=6)[0]
    mask, chunk_run, nzrow, (nruns, quiet, steps) = flat_plan(off, 0)
    nnz = int(off[-1])
    assert sum(bin(int(w)).count("1") for w in mask) == nruns == int(chunk_run[-1])
    assert nzrow[0] == -1 and nzrow[-1] == off.size - 1 and np.all(np.diff(nzrow) > 0)
    assert (int(mask[(nnz - 1) >> 5]) >> ((nnz - 1) & 31)) & 1          # the last non-zero always ends a row
    assert 0 < quiet < steps
