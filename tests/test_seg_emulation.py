"""CPU: lane-by-lane emulation of the row bookkeeping of b200::csr_seg_kernel (cudalibrarysamples_b200/csrc/spmv_csr.cu,
register path for row-sparse tiles).

What can be pinned without a GPU is the ALGORITHM: which warp owns which row end, the window of the next 32 row ends,
the end-lane mask, first-row butterfly + segmented scan, the deferred first row of every chunk and how the per-warp
partials are stitched together.  The emulator follows the kernel statement by statement (same variable names) with numpy
arrays standing in for the 32 lanes, and is checked against the oracle on matrices built to hit every branch.  Tiles that
are not row-sparse are computed by a direct restatement of tile_phase2's semantics (complete rows + head / tail
partials).  The kernel itself runs under tests/test_parity_gpu.py (csr_kernel fixture: "seg") on the GPU."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.partition_ref import csr_partition, split_rows

TILE, LONG = 2048, 512
NWARPS = 8
BIG = 32767


def warp_count_ended(S, nr, lo):
    """statement-by-statement copy of warp_count_ended()"""
    if nr <= 0:
        return 0
    lane = np.arange(32)
    stride = (nr + 31) >> 5
    blk_lo = lane * stride
    blk_hi = np.minimum(blk_lo + stride, nr)
    full = (blk_lo < nr) & (S[np.minimum(blk_hi, nr)] <= lo)
    nb = int(np.count_nonzero(full))
    cnt = nb * stride
    if cnt >= nr:
        return nr
    hi = min(cnt + stride, nr)
    extra = 0
    for j0 in range(cnt, hi, 32):
        j = j0 + lane
        extra += int(np.count_nonzero((j < hi) & (S[np.minimum(j + 1, nr)] <= lo)))
    return cnt + extra


def axpby(alpha, s, beta, yv):
    return alpha * s if beta == 0 else alpha * s + beta * yv


def emulate_tile_register_path(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b, counters):
    rs, ns = int(st[0]), int(st[1])
    re, ne = int(en[0]), int(en[1])
    lane = np.arange(32)
    nr, cnt = re - rs, ne - ns
    noff = nr + 1
    sOff = np.concatenate([np.maximum(off[rs:rs + noff] - ns, -1), np.full(34, BIG)]).astype(np.int64)   # to_soff + sentinels
    al = ns & ~31
    lead, span = ns - al, ne - al
    prod = val[ns:ne] * x[col[ns:ne]]                        # tile-relative products
    steps_total = (span + 31) >> 5
    steps_w = (steps_total + NWARPS - 1) // NWARPS
    sFirst, sOpen = np.zeros(NWARPS), np.zeros(NWARPS)
    sFirstRow = np.full(NWARPS, -1)
    for warp in range(NWARPS):
        s0 = warp * steps_w
        cs = max(s0 * 32 - lead, 0)
        ce = min((s0 + steps_w) * 32 - lead, cnt)
        active = cs < ce
        cur, frow = 0, -1
        acc, first, acc_live = np.zeros(32), 0.0, False
        if active:
            cur = 0 if warp == 0 else warp_count_ended(sOff, nr, cs)
            ws, we = sOff[cur + lane], sOff[cur + lane + 1]
            for k in range(steps_w):
                sb = (s0 + k) * 32 - lead
                if not (sb < ce):
                    continue
                pos = sb + lane
                live = (pos >= 0) & (pos < cnt)
                p = np.where(live, prod[np.clip(pos, 0, max(cnt - 1, 0))], 0.0)
                step_end = min(sb + 32, ce)
                pend = p.copy()
                while True:
                    ended = we <= step_end
                    kk = int(np.count_nonzero(ended))
                    assert np.all(ended[:kk]) and not np.any(ended[kk:])       # the ballot is a prefix
                    if kk == 0:
                        break
                    has = (lane < kk) & (we > np.maximum(ws, cs))
                    sh = np.where(has, we - 1 - sb, 0)
                    assert np.all((sh >= 0) & (sh < 32))
                    m = int(np.bitwise_or.reduce(np.where(has, 1 << sh, 0)))
                    res = np.zeros(32)
                    if m != 0:
                        counters["flush"] += 1
                        e1 = (m & -m).bit_length() - 1
                        ek = m.bit_length() - 1
                        t1 = float((acc + np.where(lane <= e1, pend, 0.0)).sum())
                        q = np.where((lane > e1) & (lane <= ek), pend, 0.0)
                        if m & (m - 1):
                            below = np.array([m & ((1 << int(l)) - 1) for l in lane])
                            dist = np.where((lane > e1) & (lane <= ek), lane - np.array([int(v).bit_length() for v in below]), 0)
                            d = 1
                            while d < 32:
                                if not np.any(dist >= d):
                                    break
                                counters["scan_levels"] += 1
                                t = np.concatenate([q[:d], q[:-d]])               # shfl_up: lanes < d keep their own value
                                q = np.where(dist >= d, q + t, q)
                                d <<= 1
                        res = np.where(lane == e1, t1, q)
                        acc = np.zeros(32)
                        acc_live = False
                        pend = np.where(lane > ek, pend, 0.0)
                    valj = np.where(has, res[sh], 0.0)
                    if frow < 0:
                        first = float(valj[0])
                        frow = cur
                        rows_out = [j for j in range(1, kk)]
                    else:
                        rows_out = [j for j in range(kk)]
                    for j in rows_out:
                        r = rs + cur + j
                        assert counters["written"][r] == 0, ("row written twice", r)
                        counters["written"][r] += 1
                        y[r] = axpby(alpha, float(valj[j]), beta, y[r])
                    cur += kk
                    ws, we = sOff[cur + lane], sOff[cur + lane + 1]
                    if kk < 32:
                        break
                acc = acc + pend
                acc_live = acc_live or bool(np.any(pend != 0))
        open_ = float(acc.sum()) if acc_live else 0.0
        if frow >= 0:
            sFirst[warp], sOpen[warp] = first, open_
        else:
            sFirst[warp], sOpen[warp] = open_, 0.0
        sFirstRow[warp] = frow
    head = sOff[0] < 0
    running, any_end = 0.0, False
    for w in range(NWARPS):
        fr = sFirstRow[w]
        if fr >= 0:
            tot = running + sFirst[w]
            if fr == 0 and head:
                head_part[b] = tot
            else:
                r = rs + fr
                assert counters["written"][r] == 0, ("row written twice", r)
                counters["written"][r] += 1
                y[r] = axpby(alpha, tot, beta, y[r])
            running = sOpen[w]
            any_end = True
        else:
            running += sFirst[w]
    if head and not any_end:
        head_part[b] = running
    elif re < rows and cnt > sOff[nr]:
        tail_part[b] = running
    else:
        assert running == 0.0


def restate_tile_phase2(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b, counters):
    """tile_phase2's semantics for the other tiles: complete rows, head partial, tail partial."""
    rs, ns = int(st[0]), int(st[1])
    re, ne = int(en[0]), int(en[1])
    prod = val[ns:ne] * x[col[ns:ne]]
    head = rs < rows and ns > off[rs]
    head_end = (min(int(off[rs + 1]), ne) - ns) if head else 0
    r_first = rs + (1 if head else 0)
    tail, tail_beg = False, ne - ns
    if re < rows and re >= r_first and ne > off[re]:
        tail, tail_beg = True, int(off[re]) - ns
    for r in range(r_first, re):
        t = float(prod[off[r] - ns:off[r + 1] - ns].sum())
        counters["written"][r] += 1
        y[r] = axpby(alpha, t, beta, y[r])
    if head:
        head_part[b] = float(prod[:head_end].sum())
    if tail:
        tail_part[b] = float(prod[tail_beg:].sum())


def emulate_spmv(off, col, val, x, y0, alpha, beta, seg_dense=24):
    off = off.astype(np.int64)
    rows = off.size - 1
    tiles = csr_partition(off, 0, TILE, LONG).astype(np.int64)
    nt = tiles.shape[0] - 1
    y = y0.astype(np.float64).copy()
    head_part, tail_part = np.full(nt + 1, np.nan), np.full(nt + 1, np.nan)      # NaN: a partial that is read must have been written
    counters = dict(flush=0, scan_levels=0, written=np.zeros(rows, int))
    n_sparse = 0
    for b in range(nt):
        st, en = tiles[b], tiles[b + 1]
        cnt, nr = int(en[1] - st[1]), int(en[0] - st[0])
        if cnt >= seg_dense * max(nr, 1):
            n_sparse += 1
            emulate_tile_register_path(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b, counters)
        else:
            restate_tile_phase2(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b, counters)
    split = split_rows(tiles, off, 0, TILE)
    for (r, b1, b2) in split:          # csr_fixup_kernel / sum_split_rows
        tot = tail_part[b1] + head_part[b1 + 1:b2 + 1].sum()
        assert counters["written"][r] == 0, ("split row written by a tile", r)
        counters["written"][r] += 1
        y[r] = axpby(alpha, tot, beta, y0[r])
    assert np.all(counters["written"] == 1), "every row is written exactly once"
    return y, n_sparse, nt, counters


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
    "rows_of_exactly_64": [64] * 70,
    "rows_of_exactly_32": [32] * 200,             # every step ends exactly one row at lane 31
    "rows_of_exactly_256": [256] * 20,            # row ends coincide with warp-chunk borders
    "rows_of_exactly_2048": [2048] * 4,           # row ends coincide with tile borders
    "dense_then_sparse": [3000, 2500, 800] + [3] * 400 + [1500, 0, 0, 900],
    "long_rows_with_empty_rows_between": [300, 0, 0, 0, 280, 0, 1000, 0, 0, 2100, 0, 70],
    "just_below_long": [511] * 9,
    "chunk_border_plus_minus_one": [255, 257, 256, 1, 255, 512, 513, 31, 33, 32, 992],
    "rmat_like_block": [500, 158, 158, 50, 158, 50, 50, 16, 158, 50, 50, 16, 50, 16, 16, 5] * 6,
    "many_rows_end_in_one_step": [1800] + [1] * 40 + [0] * 50 + [2] * 30 + [1900, 0, 0, 0, 1, 1, 1],
    "long_row_then_40_empty_rows": [2500] + [0] * 40 + [1500] + [0] * 70 + [30, 2000],
    "split_row_ends_on_tile_border": [100, 1948 + 2048, 5, 2048 * 3 - 2053, 7],
}


@pytest.mark.parametrize("name", list(PROFILES))
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-2.0, 0.5)])
@pytest.mark.parametrize("seg_dense", [1, 24])
def test_register_path_bookkeeping(name, alpha, beta, seg_dense):
    lens = np.array(PROFILES[name])
    cols = 30000
    off, col, val = build(lens, cols, 7)
    x = O.uniform(1, cols)
    y0 = O.uniform(2, lens.size)
    want = O.spmv_csr(off, col, val, x, y0, alpha, beta)
    got, n_sparse, nt, _ = emulate_spmv(off, col, val, x, y0, alpha, beta, seg_dense)
    assert n_sparse > 0, "profile must exercise the register path"
    assert np.linalg.norm(got - want) <= 1e-12 * max(np.linalg.norm(want), 1e-300)


@pytest.mark.parametrize("seed", range(6))
def test_register_path_random_structures(seed):
    """random mixtures of empty, short, medium and very long rows; seg_dense = 1 sends EVERY tile with at least one
    non-zero per row end down the register path, so the multi-row scan and the >32-rows-per-step loop get exercised"""
    rng = np.random.default_rng(100 + seed)
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
    got, n_sparse, nt, _ = emulate_spmv(off, col, val, x, y0, 0.75, 1.25, seg_dense=1)
    assert n_sparse > 0
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


def test_register_path_on_rmat():
    off, col, val = O.rmat_csr(40000, avg_nnz=16, seed=3, val_seed=4)
    x, y0 = O.uniform(5, 40000), O.uniform(6, 40000)
    want = O.spmv_csr(off, col, val, x, y0, 1.5, -0.25)
    got, n_sparse, nt, c = emulate_spmv(off, col, val, x, y0, 1.5, -0.25)
    assert 0 < n_sparse < nt            # R-MAT has both kinds of tiles
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
