"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/liboracle.so (built from oracle/spmv_oracle.c by oracle/Makefile) plus the
numpy host logic that turns the oracle's R-MAT edge stream into a CSR matrix.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg.

Reference call sites restated here (paths relative to /root/reference):
  * y = alpha*A*x + beta*y ...... cuSPARSE/spmvop_csr/spmv_csr_op_example.c:307-318
  * 4x4 toy problem ............. cuSPARSE/spmv_csr/spmv_csr_example.c:45-56 (+ spmv_coo, spmv_sell)
  * generators .................. cuSPARSE/cg/cg_example.c:71-128,
                                  cuSPARSE/bicgstab/bicgstab_example.c:69-127,
                                  cuDSS/simple_residual/laplace_generator.hxx:34-107
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RMAT_ABCD = (0.57, 0.19, 0.19, 0.05)


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (seconds).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spmv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_time_csr_f64.restype = C.c_double
        for n in ("oracle_gen_stencil5", "oracle_gen_laplace7", "oracle_csr_to_sell_f64", "oracle_csr_to_sell_f32"):
            getattr(_LIB, n).restype = C.c_int64
    return _LIB


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


# --------------------------------------------------------------------------------------------------
# y = alpha*A*x + beta*y
# --------------------------------------------------------------------------------------------------
def spmv_csr(off, col, val, x, y=None, alpha=1.0, beta=0.0, base=0, threads=1):
    off = np.ascontiguousarray(off, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    x = np.ascontiguousarray(x, val.dtype)
    rows = off.size - 1
    out = np.zeros(rows, val.dtype) if y is None else np.array(y, val.dtype, copy=True)
    getattr(lib(), "oracle_spmv_csr_" + _vt(val.dtype))(
        C.c_int64(rows), _p(off), _p(col), _p(val), C.c_int32(base), C.c_double(alpha), C.c_double(beta),
        _p(x), _p(out), C.c_int(threads))
    return out


def spmm_csr(off, col, val, B, C0=None, alpha=1.0, beta=0.0, base=0, order_b="col", order_c="col", threads=1):
    """C = alpha*A*B + beta*C0 with B (cols x n) and C0 (rows x n) given as 2-D numpy arrays; `order_*` = memory layout the
    C loop walks ("col": spmm_csr_example.c:50-66, "row": spmm_csr_op_example.c:195-199).  Returns C as a 2-D array."""
    off = np.ascontiguousarray(off, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    rows, n = off.size - 1, B.shape[1]
    Bm = np.ascontiguousarray(B, val.dtype) if order_b == "row" else np.asfortranarray(B, val.dtype)
    out = np.zeros((rows, n), val.dtype) if C0 is None else np.array(C0, val.dtype)
    Cm = np.ascontiguousarray(out) if order_c == "row" else np.asfortranarray(out)
    ldb = n if order_b == "row" else B.shape[0]
    ldc = n if order_c == "row" else rows
    getattr(lib(), "oracle_spmm_csr_" + _vt(val.dtype))(
        C.c_int64(rows), C.c_int64(n), _p(off), _p(col), _p(val), C.c_int32(base), C.c_double(alpha), C.c_double(beta), _p(Bm),
        C.c_int64(ldb), C.c_int(order_b == "row"), _p(Cm), C.c_int64(ldc), C.c_int(order_c == "row"), C.c_int(threads))
    return np.array(Cm)


def spmv_coo(rows, row, col, val, x, y=None, alpha=1.0, beta=0.0, base=0):
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    x = np.ascontiguousarray(x, val.dtype)
    out = np.zeros(rows, val.dtype) if y is None else np.array(y, val.dtype, copy=True)
    getattr(lib(), "oracle_spmv_coo_" + _vt(val.dtype))(
        C.c_int64(rows), C.c_int64(row.size), _p(row), _p(col), _p(val), C.c_int32(base), C.c_double(alpha),
        C.c_double(beta), _p(x), _p(out))
    return out


def spmv_sell(rows, slice_size, slice_off, col, val, x, y=None, alpha=1.0, beta=0.0, base=0, threads=1):
    slice_off = np.ascontiguousarray(slice_off, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    x = np.ascontiguousarray(x, val.dtype)
    out = np.zeros(rows, val.dtype) if y is None else np.array(y, val.dtype, copy=True)
    getattr(lib(), "oracle_spmv_sell_" + _vt(val.dtype))(
        C.c_int64(rows), C.c_int64(slice_size), _p(slice_off), _p(col), _p(val), C.c_int32(base),
        C.c_double(alpha), C.c_double(beta), _p(x), _p(out), C.c_int(threads))
    return out


def time_csr_f64(off, col, val, x, threads, reps=3, alpha=1.0, beta=0.0):
    """Wall-clock seconds (min over reps, and the list) of the CSR loop on `threads` host threads."""
    rows = off.size - 1
    y = np.zeros(rows, np.float64)
    times = np.zeros(reps, np.float64)
    best = lib().oracle_time_csr_f64(C.c_int64(rows), _p(off), _p(col), _p(val), C.c_double(alpha),
                                     C.c_double(beta), _p(x), _p(y), C.c_int(threads), C.c_int(reps), _p(times))
    return float(best), times.tolist(), y


# --------------------------------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------------------------------
def gen_stencil5(grid, mass=0.04, ux=0.0, uy=0.0):
    """cg_example.c:71-128 (defaults) / bicgstab_example.c:69-127 (mass=.3, ux=.3, uy=.2)."""
    L = lib()
    n = grid * grid
    nnz = L.oracle_gen_stencil5(C.c_int32(grid), C.c_double(mass), C.c_double(ux), C.c_double(uy), None, None, None)
    off = np.empty(n + 1, np.int32)
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.float64)
    got = L.oracle_gen_stencil5(C.c_int32(grid), C.c_double(mass), C.c_double(ux), C.c_double(uy), _p(off), _p(col), _p(val))
    assert got == nnz
    return off, col, val


def gen_laplace7(nx):
    """cuDSS/simple_residual/laplace_generator.hxx:34-107."""
    L = lib()
    n = nx ** 3
    nnz = L.oracle_gen_laplace7(C.c_int32(nx), None, None, None)
    off = np.empty(n + 1, np.int32)
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.float64)
    got = L.oracle_gen_laplace7(C.c_int32(nx), _p(off), _p(col), _p(val))
    assert got == nnz
    return off, col, val


def rmat_thresholds(abcd=RMAT_ABCD):
    a, b, c, _ = abcd
    two32 = 4294967296.0
    return int(a * two32), int((a + b) * two32), int((a + b + c) * two32)


def rmat_edges(seed, e0, count, scale, abcd=RMAT_ABCD):
    r = np.empty(count, np.int64)
    c = np.empty(count, np.int64)
    tA, tAB, tABC = rmat_thresholds(abcd)
    lib().oracle_rmat_edges(C.c_uint64(seed), C.c_int64(e0), C.c_int64(count), C.c_int32(scale),
                            C.c_uint64(tA), C.c_uint64(tAB), C.c_uint64(tABC), _p(r), _p(c))
    return r, c


def uniform(seed, count, dtype=np.float64, i0=0):
    out = np.empty(count, dtype)
    getattr(lib(), "oracle_uniform_" + _vt(dtype))(C.c_uint64(seed), C.c_int64(i0), C.c_int64(count), _p(out))
    return out


def rmat_scale(n):
    return max(1, int(n - 1).bit_length())


RMAT_OVERSAMPLE = 1.5  # candidate edges generated per wanted non-zero (rejection + duplicates)


def rmat_csr(rows, cols=None, avg_nnz=16, seed=42, val_seed=43, dtype=np.float64, abcd=RMAT_ABCD):
    """R-MAT, rejection to rows x cols, duplicates merged, columns sorted.

    Definition (shared bit for bit with cudalibrarysamples_b200.workloads.rmat_csr on the GPU):
    candidate edges e = 0 .. ceil(1.5*rows*avg_nnz)-1 from the hash stream; drop edges outside
    rows x cols; keep the FIRST occurrence of each (row, col) in stream order; keep the first
    rows*avg_nnz of those; sort by (row, col).  val[j] = U(-1,1) from hash(val_seed, j) in CSR order.
    """
    cols = rows if cols is None else cols
    scale = rmat_scale(max(rows, cols))
    target = int(rows) * int(avg_nnz)
    cand = int(np.ceil(target * RMAT_OVERSAMPLE))
    r, c = rmat_edges(seed, 0, cand, scale, abcd)
    ok = (r < rows) & (c < cols)
    key = r[ok] * np.int64(cols) + c[ok]
    _, first = np.unique(key, return_index=True)
    first.sort()
    keep = np.sort(key[first[:target]])
    rr = keep // cols
    cc = (keep % cols).astype(np.int32)
    off = np.zeros(rows + 1, np.int64)
    off[1:] = np.bincount(rr, minlength=rows)
    off = np.cumsum(off).astype(np.int32)
    val = uniform(val_seed, keep.size, dtype)
    return off, cc, val


# --------------------------------------------------------------------------------------------------
# converters
# --------------------------------------------------------------------------------------------------
def csr_to_coo_rows(off, base=0):
    off = np.ascontiguousarray(off, np.int32)
    rows = off.size - 1
    out = np.empty(int(off[-1]) - base, np.int32)
    lib().oracle_csr_to_coo_rows(C.c_int64(rows), _p(off), C.c_int32(base), _p(out))
    return out


def csr_to_sell(off, col, val, slice_size, base=0):
    off = np.ascontiguousarray(off, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    rows = off.size - 1
    nslices = (rows + slice_size - 1) // slice_size
    so = np.empty(nslices + 1, np.int32)
    fn = getattr(lib(), "oracle_csr_to_sell_" + _vt(val.dtype))
    size = fn(C.c_int64(rows), _p(off), _p(col), _p(val), C.c_int32(base), C.c_int64(slice_size), _p(so), None, None)
    co = np.empty(size, np.int32)
    vo = np.empty(size, val.dtype)
    fn(C.c_int64(rows), _p(off), _p(col), _p(val), C.c_int32(base), C.c_int64(slice_size), _p(so), _p(co), _p(vo))
    return so, co, vo


# The reference's toy problem (spmv_csr_example.c:45-56; same matrix in spmv_coo / spmv_sell).
TOY = dict(
    rows=4, cols=4, nnz=9,
    csr_off=np.array([0, 3, 4, 7, 9], np.int32),
    csr_col=np.array([0, 2, 3, 1, 0, 2, 3, 1, 3], np.int32),
    val=np.array([1, 2, 3, 4, 5, 6, 7, 8, 9], np.float32),
    coo_row=np.array([0, 0, 0, 1, 2, 2, 2, 3, 3], np.int32),                  # spmv_coo_example.c:48
    sell_slice_size=2, sell_values_size=12,
    sell_off=np.array([0, 6, 12], np.int32),                                   # spmv_sell_example.c:52
    sell_col=np.array([0, 1, 2, -1, 3, -1, 0, 1, 2, 3, 3, -1], np.int32),      # spmv_sell_example.c:53-60
    sell_val=np.array([1, 4, 2, 0, 3, 0, 5, 8, 6, 9, 7, 0], np.float32),       # spmv_sell_example.c:61-66
    x=np.array([1, 2, 3, 4], np.float32),
    y_result=np.array([19, 8, 51, 52], np.float32),                            # spmv_csr_example.c:54
    # spmm_csr_example.c:59-66: B 4x3 and the golden C 4x3, both column-major
    spmm_B=np.arange(1, 13, dtype=np.float32).reshape(3, 4).T.copy(),
    spmm_C=np.array([19, 8, 51, 52, 43, 24, 123, 120, 67, 40, 195, 188], np.float32).reshape(3, 4).T.copy(),
)

# The strided-batch variant of the toy product (cuSPARSE/spmm_csr_batched/spmm_csr_batched_example.c:56-88): two 4x4 matrices
# sharing the row offsets (:56), columns / values per batch (:57-67), B per batch (:68-73), golden C per batch (:80-85), all
# dense operands column-major 4x3; the sample compares with `!=` (:183-196).
TOY_BATCHED = dict(
    rows=4, cols=4, nnz=9, n=3, batches=2,
    csr_off=np.array([0, 3, 4, 7, 9], np.int32),
    csr_col=np.array([[0, 2, 3, 1, 0, 2, 3, 1, 3], [1, 2, 3, 0, 0, 1, 3, 1, 2]], np.int32),
    val=np.array([[1, 2, 3, 4, 5, 6, 7, 8, 9], [10, 11, 12, 13, 14, 15, 16, 17, 18]], np.float32),
    B=np.array([[1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], [6, 4, 3, 2, 1, 6, 9, 8, 9, 3, 2, 5]], np.float32),          # column-major buffers
    C=np.array([[19, 8, 51, 52, 43, 24, 123, 120, 67, 40, 195, 188], [97, 78, 176, 122, 255, 13, 232, 264, 112, 117, 251, 87]], np.float32),
)
