"""oracle/partition_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the integer preprocessing that cusparseSpMV_preprocess performs in this build
(b200::csr_partition_kernel in cudalibrarysamples_b200/csrc/spmv_csr.cu): the tile partition of the CSR
(rows + nnz) merge list.  The reference's own preprocessing (cusparse::partition_kernel in the closed
libcusparse) is not observable, so parity for this integer work is defined against this restatement:
tests/test_parity_gpu.py demands bit-exact equality of every (row, nnz) tile coordinate.
"""
from __future__ import annotations

import numpy as np


def csr_partition(off, base, tile_items, long_row):
    """Returns int32 array [num_tiles+1, 2] of (row, nnz) tile start coordinates.

    boundary b sits on merge diagonal d = min(b*tile_items, rows+nnz); r = largest row index with
    r + off[r] <= d.  Rows shorter than long_row are never cut (boundary = row start); longer rows are cut
    exactly at the diagonal.
    """
    off = np.asarray(off, np.int64) - base
    rows = off.size - 1
    nnz = int(off[-1])
    num_tiles = (rows + nnz + tile_items - 1) // tile_items
    d = np.minimum(np.arange(num_tiles + 1, dtype=np.int64) * tile_items, rows + nnz)
    g = np.arange(rows + 1, dtype=np.int64) + off          # strictly increasing
    r = np.searchsorted(g, d, side="right") - 1            # largest r with g[r] <= d
    n = off[r].copy()
    inside = r < rows
    rr = np.where(inside, r, 0)
    length = np.where(inside, off[np.minimum(rr + 1, rows)] - off[rr], 0)
    e = d - (r + n)
    cut = inside & (length >= long_row) & (e > 0)
    n = np.where(cut, n + e, n)
    return np.stack([r, n], axis=1).astype(np.int32)


def check_partition(tiles, off, base, tile_items, long_row):
    """Invariants every valid partition must satisfy (used on CPU and at full size on the GPU)."""
    off = np.asarray(off, np.int64) - base
    rows = off.size - 1
    nnz = int(off[-1])
    t = tiles.astype(np.int64)
    assert t[0, 0] == 0 and t[0, 1] == 0
    assert t[-1, 0] == rows and t[-1, 1] == nnz
    pos = t[:, 0] + t[:, 1]
    assert np.all(np.diff(pos) >= 0)
    assert np.all(np.diff(t[:, 0]) >= 0) and np.all(np.diff(t[:, 1]) >= 0)
    # a tile never holds more items than shared memory was sized for
    assert np.all(np.diff(pos) <= tile_items + long_row - 1)
    # a boundary is either a row start or strictly inside a long row
    r, n = t[:, 0], t[:, 1]
    rs = np.minimum(r, rows)
    start = off[rs]
    mid = n != start
    if mid.any():
        rl = r[mid]
        assert np.all(rl < rows)
        assert np.all(off[rl + 1] - off[rl] >= long_row)
        assert np.all((n[mid] > off[rl]) & (n[mid] <= off[rl + 1]))
    return True


def split_rows(tiles, off, base, tile_items):
    """Rows cut by a tile boundary, as (row, first covering tile b1, last covering tile b2), sorted by row.

    Restates the split list that csr_partition_kernel registers: boundary b lies strictly inside row r
    (tiles[b].nnz > off[r]); b1 = g(r) // tile_items, b2 = (g(r) + len(r)) // tile_items with g(r) = r + off[r];
    the row is registered by its first cut, b == b1 + 1.
    """
    off = np.asarray(off, np.int64) - base
    rows = off.size - 1
    t = tiles.astype(np.int64)
    r, n = t[:, 0], t[:, 1]
    mid = (r < rows) & (n > off[np.minimum(r, rows)])
    out = []
    for b in np.nonzero(mid)[0]:
        rr = int(r[b])
        g0 = rr + int(off[rr])
        b1 = g0 // tile_items
        b2 = (g0 + int(off[rr + 1] - off[rr])) // tile_items
        if b == b1 + 1:
            out.append((rr, b1, b2))
    return sorted(out)


# ---------------------------------------------------------------------------------------------------------------------
# flat plan (b200spmv_csr_flat_analyze in cudalibrarysamples_b200/csrc/spmv_csr_flat.cu): integer preprocessing, bit-exact
# ---------------------------------------------------------------------------------------------------------------------
FLAT_CHUNK = 256          # non-zeros per warp chunk
FLAT_CTA_NNZ = 2048       # non-zeros per CTA


def flat_plan(off, base):
    """numpy restatement of the flat CSR plan:
       endmask[w] bit i  <=> non-zero 32*w+i is the last one of its row (zero-padded to whole CTAs of 2048 non-zeros),
       chunk_run[c]      =   number of rows that end before non-zero 256*c   (c = 0 .. nchunks),
       nzrow             =   [-1] + [indices of the non-empty rows] + [rows],
       ctl               =   (non-empty rows, 32-non-zero steps in which no row ends, steps)."""
    off = np.asarray(off, np.int64) - base
    rows = off.size - 1
    nnz = int(off[-1])
    nz = np.diff(off) > 0
    ends = off[1:][nz] - 1
    nctas = (nnz + FLAT_CTA_NNZ - 1) // FLAT_CTA_NNZ
    mask = np.zeros(nctas * (FLAT_CTA_NNZ // 32), np.uint32)
    np.bitwise_or.at(mask, ends >> 5, (np.uint32(1) << (ends & 31).astype(np.uint32)))
    nchunks = nctas * (FLAT_CTA_NNZ // FLAT_CHUNK)
    per_chunk = np.bincount(ends // FLAT_CHUNK, minlength=nchunks)[:nchunks] if nchunks else np.zeros(0, np.int64)
    chunk_run = np.concatenate([[0], np.cumsum(per_chunk)]).astype(np.int32)
    nzrow = np.concatenate([[-1], np.nonzero(nz)[0], [rows]]).astype(np.int32)
    steps = (nnz + 31) // 32
    quiet = int(np.count_nonzero(mask[:steps] == 0))
    return mask, chunk_run, nzrow, (int(nz.sum()), quiet, steps)
