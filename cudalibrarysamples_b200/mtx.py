"""Matrix Market ingestion for the SpMV path: host-side mirror of the reference's reader
cuDSS/simple_matrix_market/matrix_market_reader.h:36-191 (`matrix_reader`), so real (SuiteSparse) matrices can be run
through the same cusparseCreateCsr / cusparseSpMV sequence as the synthetic ones.

Behaviour kept from the reference reader: header must be `%%MatrixMarket matrix coordinate real|integer|pattern ...`
(:74-91; the reference accepts only `real`), comment lines start with `%` (:70-71), 1-based indices (:107-108), the entry
count must match the size line (:133-137 -> MtxReaderErrorWrongNnz), indices outside the matrix are errors (:164-171),
entries are sorted by (row, column) (:139) and turned into CSR offsets by counting + prefix sum (:172-179); empty rows are
legal (:182-187).  Added: `symmetric` / `skew-symmetric` files can be expanded to the full matrix (the reference reader
leaves that to cuDSS's matrix view), rectangular sizes, `pattern` files get value 1.
"""
from __future__ import annotations

import numpy as np


class MtxReaderError(ValueError):
    """Mirrors enum MtxReaderStatus (matrix_market_reader.h:24-34); `.status` holds the reference's name."""

    def __init__(self, status: str, msg: str):
        super().__init__(f"{status}: {msg}")
        self.status = status


def read_matrix_market(path: str, expand_symmetric: bool = True, dtype=np.float64):
    """Returns (rows, cols, csr_offsets int32[rows+1], csr_columns int32[nnz], csr_values dtype[nnz]), base 0."""
    try:
        f = open(path, "r")
    except OSError as e:
        raise MtxReaderError("MtxReaderErrorFileNotFound", str(e))
    with f:
        header = None
        size = None
        rows_i, cols_i, vals = [], [], []
        for line in f:
            line = line.strip()
            if not line:
                continue
            if header is None and line.startswith("%%MatrixMarket"):
                parts = line.split()
                header = [p.lower() for p in parts[1:]]
                if len(header) < 4 or header[0] != "matrix" or header[1] != "coordinate" or header[2] not in ("real", "integer", "pattern"):
                    raise MtxReaderError("MtxReaderErrorInvalidFormatInHeader", line)
                continue
            if line[0] == "%":
                continue
            tok = line.split()
            if size is None:
                if len(tok) < 3:
                    raise MtxReaderError("MtxReaderErrorInvalidFormatInHeader", "size line: " + line)
                size = (int(tok[0]), int(tok[1]), int(tok[2]))
                continue
            rows_i.append(int(tok[0]) - 1)
            cols_i.append(int(tok[1]) - 1)
            vals.append(float(tok[2]) if len(tok) > 2 else 1.0)
    if header is None or size is None:
        raise MtxReaderError("MtxReaderErrorInvalidFormatInHeader", "no %%MatrixMarket header / size line")
    n_rows, n_cols, declared = size
    if len(vals) != declared:
        raise MtxReaderError("MtxReaderErrorWrongNnz", f"{len(vals)} entries in the file, {declared} announced")
    r = np.asarray(rows_i, np.int64)
    c = np.asarray(cols_i, np.int64)
    v = np.asarray(vals, dtype)
    if r.size and (r.min() < 0 or r.max() >= n_rows):
        raise MtxReaderError("MtxReaderErrorOutOfBoundRowIndex", "row index outside the matrix")
    if c.size and (c.min() < 0 or c.max() >= n_cols):
        raise MtxReaderError("MtxReaderErrorOfBoundColIndex", "column index outside the matrix")
    symmetry = header[3] if len(header) > 3 else "general"
    if expand_symmetric and symmetry in ("symmetric", "skew-symmetric", "hermitian"):
        offd = r != c
        sign = -1.0 if symmetry == "skew-symmetric" else 1.0
        r, c, v = np.concatenate([r, c[offd]]), np.concatenate([c, r[offd]]), np.concatenate([v, sign * v[offd]])
    order = np.lexsort((c, r))                                # sorted by (row, column), like std::sort on the tuples
    r, c, v = r[order], c[order], v[order]
    off = np.zeros(n_rows + 1, np.int64)
    np.add.at(off, r + 1, 1)
    off = np.cumsum(off).astype(np.int32)
    return n_rows, n_cols, off, c.astype(np.int32), v


def write_matrix_market(path: str, rows: int, cols: int, off, col, val, comment: str = ""):
    """`%%MatrixMarket matrix coordinate real general`, 1-based, one entry per line, 17 significant digits."""
    off = np.asarray(off)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        if comment:
            f.write("% " + comment + "\n")
        f.write(f"{rows} {cols} {int(off[-1] - off[0])}\n")
        for i in range(rows):
            for p in range(int(off[i] - off[0]), int(off[i + 1] - off[0])):
                f.write(f"{i + 1} {int(col[p]) + 1} {float(val[p])!r}\n")
