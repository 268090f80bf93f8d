"""CPU: lane-by-lane emulation of the row bookkeeping of the round-2 candidate kernel b200::csr_hyb_kernel
(cudalibrarysamples_b200/csrc/spmv_csr.cu, register path for tiles with long rows).

The CUDA kernel has not run on hardware yet; what can be pinned without a GPU is its ALGORITHM: which warp owns which
row end, when a row counts as started / open / first, and how the per-warp partials are stitched together.  The
emulator below follows the kernel statement by statement (same variable names) with numpy arrays standing in for the
32 lanes, and is checked against the oracle on matrices built to hit every branch.  Tiles that are not "dense" are
computed by a direct restatement of tile_phase2's semantics (complete rows + head / tail partials)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle.partition_ref import csr_partition, split_rows

TILE, LONG, DENSE = 2048, 512, 64
HAS_FIRST, FIRST_ENDS, HAS_OPEN = 1, 2, 4
INF = 0x7FFFFFFF


def emulate_tile_register_path(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b):
    rs, ns = int(st[0]), int(st[1])
    re, ne = int(en[0]), int(en[1])
    lane = np.arange(32)
    noff = min(re, rows) - rs + 1
    sOff = np.maximum(off[rs:rs + noff] - ns, -1)            # to_soff
    al = ns & ~31
    lead, span, cnt = ns - al, ne - al, ne - ns
    nr = re - rs
    prod = val[ns:ne] * x[col[ns:ne]]                        # tile-relative products
    steps_total = (span + 31) >> 5
    steps_w = (steps_total + 7) >> 3
    sFirst, sOpen = np.zeros(8), np.zeros(8)
    sOpenRow, sFlags = np.zeros(8, int), np.zeros(8, int)
    for warp in range(8):
        s0 = warp * steps_w
        cs = max(s0 * 32 - lead, 0)
        ce = min((s0 + steps_w) * 32 - lead, cnt)
        active = cs < ce
        cur, seg_beg, cur_end, flags, open_row = 0, cs, INF, 0, 0
        started, acc, first, open_ = True, np.zeros(32), 0.0, 0.0
        if active:
            lo = -1 if warp == 0 else cs
            n = 0
            for r in range(2):
                j = lane + 1 + 32 * r
                ok = (j <= nr) & (sOff[np.minimum(j, noff - 1)] <= lo)
                n += int(np.count_nonzero(ok))
            cur = n
            s_cur = int(sOff[cur])
            started = (s_cur >= 0) if warp == 0 else (s_cur == cs)
            cur_end = int(sOff[cur + 1]) if cur < nr else INF
            for k in range(steps_w):
                pos = (s0 + k) * 32 + lane - lead
                live = (pos >= 0) & (pos < cnt)
                p = np.where(live, prod[np.clip(pos, 0, max(cnt - 1, 0))], 0.0)
                step_end = min((s0 + k) * 32 + 32 - lead, cnt)
                while cur_end <= step_end:
                    acc = acc + np.where((pos >= seg_beg) & (pos < cur_end), p, 0.0)
                    tot = float(acc.sum())
                    if started:
                        r = rs + cur
                        y[r] = alpha * tot + (beta * y[r] if beta != 0 else 0.0)
                    else:
                        first = tot
                        flags |= HAS_FIRST | FIRST_ENDS
                    seg_beg = cur_end
                    cur += 1
                    cur_end = int(sOff[cur + 1]) if cur < nr else INF
                    started = True
                    acc = np.zeros(32)
                acc = acc + np.where((pos >= seg_beg) & (pos < step_end), p, 0.0)
            if seg_beg < ce:
                tot = float(acc.sum())
                if started:
                    open_, open_row = tot, cur
                    flags |= HAS_OPEN
                else:
                    first = tot
                    flags |= HAS_FIRST
        sFirst[warp], sOpen[warp], sOpenRow[warp], sFlags[warp] = first, open_, open_row, flags
    for w in range(8):
        f = sFlags[w]
        if f & HAS_OPEN:
            tot, ended = sOpen[w], False
            for v in range(w + 1, 8):
                fv = sFlags[v]
                if not (fv & HAS_FIRST):
                    break
                tot += sFirst[v]
                if fv & FIRST_ENDS:
                    ended = True
                    break
            if ended:
                r = rs + sOpenRow[w]
                y[r] = alpha * tot + (beta * y[r] if beta != 0 else 0.0)
            else:
                tail_part[b] = tot
        if w == 0 and (f & HAS_FIRST):
            tot = sFirst[0]
            if not (f & FIRST_ENDS):
                for v in range(1, 8):
                    fv = sFlags[v]
                    if not (fv & HAS_FIRST):
                        break
                    tot += sFirst[v]
                    if fv & FIRST_ENDS:
                        break
            head_part[b] = tot


def restate_tile_phase2(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b):
    """tile_phase2's semantics for the non-dense tiles: complete rows, head partial, tail partial."""
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
        y[r] = alpha * t + (beta * y[r] if beta != 0 else 0.0)
    if head:
        head_part[b] = float(prod[:head_end].sum())
    if tail:
        tail_part[b] = float(prod[tail_beg:].sum())


def emulate_spmv(off, col, val, x, y0, alpha, beta):
    off = off.astype(np.int64)
    rows = off.size - 1
    tiles = csr_partition(off, 0, TILE, LONG).astype(np.int64)
    nt = tiles.shape[0] - 1
    y = y0.astype(np.float64).copy()
    head_part, tail_part = np.zeros(nt + 1), np.zeros(nt + 1)
    n_dense = 0
    for b in range(nt):
        st, en = tiles[b], tiles[b + 1]
        cnt, nr = int(en[1] - st[1]), int(en[0] - st[0])
        if cnt >= DENSE * max(nr, 1):
            n_dense += 1
            emulate_tile_register_path(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b)
        else:
            restate_tile_phase2(off, col, val, x, y, alpha, beta, rows, st, en, head_part, tail_part, b)
    for (r, b1, b2) in split_rows(tiles, off, 0, TILE):          # csr_fixup_kernel / sum_split_rows
        tot = tail_part[b1] + head_part[b1 + 1:b2 + 1].sum()
        y[r] = alpha * tot + (beta * y0[r] if beta != 0 else 0.0)
    return y, n_dense, nt


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
    "rows_of_exactly_256": [256] * 20,            # row ends coincide with warp-chunk borders
    "rows_of_exactly_2048": [2048] * 4,           # row ends coincide with tile borders
    "dense_then_sparse": [3000, 2500, 800] + [3] * 400 + [1500, 0, 0, 900],
    "long_rows_with_empty_rows_between": [300, 0, 0, 0, 280, 0, 1000, 0, 0, 2100, 0, 70],
    "just_below_long": [511] * 9,
    "chunk_border_plus_minus_one": [255, 257, 256, 1, 255, 512, 513, 31, 33, 32, 992],
}


@pytest.mark.parametrize("name", list(PROFILES))
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-2.0, 0.5)])
def test_register_path_bookkeeping(name, alpha, beta):
    lens = np.array(PROFILES[name])
    cols = 30000
    off, col, val = build(lens, cols, 7)
    x = O.uniform(1, cols)
    y0 = O.uniform(2, lens.size)
    want = O.spmv_csr(off, col, val, x, y0, alpha, beta)
    got, n_dense, nt = emulate_spmv(off, col, val, x, y0, alpha, beta)
    assert n_dense > 0, "profile must exercise the register path"
    assert np.linalg.norm(got - want) <= 1e-12 * max(np.linalg.norm(want), 1e-300)


def test_register_path_on_rmat():
    off, col, val = O.rmat_csr(40000, avg_nnz=16, seed=3, val_seed=4)
    x, y0 = O.uniform(5, 40000), O.uniform(6, 40000)
    want = O.spmv_csr(off, col, val, x, y0, 1.5, -0.25)
    got, n_dense, nt = emulate_spmv(off, col, val, x, y0, 1.5, -0.25)
    assert 0 < n_dense < nt            # R-MAT has both kinds of tiles
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
