"""Host-side mirror of the cuSPARSE generic API that the reference samples call for SpMV.

Same names, argument order and error behaviour as the C calls in cuSPARSE/spmv_csr/spmv_csr_example.c:86-118
(cusparseCreate ... cusparseCreateCsr ... cusparseSpMV_bufferSize/_preprocess/cusparseSpMV ... cusparseDestroy*), so
the parity tests read like the reference's own samples.  Every call goes through the C ABI:

  impl="b200"      descriptor + SpMV symbols come from libb200spmv.so (the product: sm_100a kernels)
  impl="cusparse"  the same symbols come from the closed libcusparse.so.12 (the GPU oracle)

Handle management (cusparseCreate/Destroy/SetStream/SetPointerMode) always belongs to the real library, exactly as
when a sample is linked with `-lb200spmv -lcusparse`.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib

# enums (cusparse.h:260-278, 4988-5007, 5669-5676; library_types.h)
CUSPARSE_STATUS_SUCCESS = 0
CUSPARSE_STATUS_INVALID_VALUE = 3
CUSPARSE_STATUS_NOT_SUPPORTED = 10
CUSPARSE_OPERATION_NON_TRANSPOSE = 0
CUSPARSE_OPERATION_TRANSPOSE = 1
CUSPARSE_INDEX_32I = 2
CUSPARSE_INDEX_64I = 3
CUSPARSE_INDEX_BASE_ZERO = 0
CUSPARSE_INDEX_BASE_ONE = 1
CUSPARSE_POINTER_MODE_HOST = 0
CUSPARSE_POINTER_MODE_DEVICE = 1
CUSPARSE_SPMV_ALG_DEFAULT = 0
CUSPARSE_SPMV_COO_ALG1 = 1
CUSPARSE_SPMV_CSR_ALG1 = 2
CUSPARSE_SPMV_CSR_ALG2 = 3
CUSPARSE_SPMV_COO_ALG2 = 4
CUSPARSE_SPMV_SELL_ALG1 = 5
CUDA_R_32F = 0
CUDA_R_64F = 1
CUSPARSE_ORDER_COL = 1
CUSPARSE_ORDER_ROW = 2
CUSPARSE_SPMM_ALG_DEFAULT = 0
CUSPARSE_SPMM_CSR_ALG2 = 6

_VT = {torch.float32: CUDA_R_32F, torch.float64: CUDA_R_64F}
_IT = {torch.int32: CUSPARSE_INDEX_32I, torch.int64: CUSPARSE_INDEX_64I}
_CT = {torch.float32: C.c_float, torch.float64: C.c_double}


class CuSparseError(RuntimeError):
    """Raised where the samples' CHECK_CUSPARSE macro would print and exit (spmv_csr_example.c:33-41)."""

    def __init__(self, fn, status):
        super().__init__(f"{fn} failed with cusparseStatus_t {status}")
        self.status = status


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


class Api:
    """One instance per implementation under test."""

    def __init__(self, impl: str = "b200", lib_path: str | None = None):
        if impl not in ("b200", "cusparse"):
            raise ValueError(impl)
        self.impl = impl
        self.real = _lib.real()
        if impl == "cusparse":
            self.lib = self.real
        elif lib_path is None:
            self.lib = _lib.shim()
        else:   # a tuning variant of the product library (scripts/sweep.py); same ABI
            _lib.shim()
            self.lib = C.CDLL(lib_path, mode=C.RTLD_LOCAL)

    # ---- run-time switches / call counters of the product library (include/b200spmv.h) --------------
    def set_option(self, key: str, value: str) -> None:
        """b200spmv_set_option: e.g. ("B200SPMV_CSR_KERNEL", "seg" | "tile" | "pipe" | "ws" | "rowwise" | "auto")."""
        if self.impl != "b200":
            raise ValueError("options belong to the b200 library")
        if self.lib.b200spmv_set_option(key.encode(), value.encode()) != 0:
            raise ValueError(f"unknown option {key}={value}")

    def stats(self) -> dict:
        """SpMV calls served by our kernels / forwarded to the closed library / CSR analyses run, since the last reset."""
        n, f, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        (self.lib if self.impl == "b200" else _lib.shim()).b200spmv_get_stats(C.byref(n), C.byref(f), C.byref(a))
        return dict(native=int(n.value), forwarded=int(f.value), analyze=int(a.value))

    def last_csr_kernel(self) -> str:
        lib = self.lib if self.impl == "b200" else _lib.shim()
        lib.b200spmv_last_csr_kernel.restype = C.c_char_p
        return lib.b200spmv_last_csr_kernel().decode()

    def reset_stats(self) -> None:
        (self.lib if self.impl == "b200" else _lib.shim()).b200spmv_reset_stats()

    # ---- helpers ---------------------------------------------------------------------------------
    def _call(self, lib, name, *args):
        st = getattr(lib, name)(*args)
        if st != CUSPARSE_STATUS_SUCCESS:
            raise CuSparseError(name, st)

    # ---- handle (always the real library) ----------------------------------------------------------
    def cusparseCreate(self):
        h = C.c_void_p()
        self._call(self.real, "cusparseCreate", C.byref(h))
        self.cusparseSetStream(h, torch.cuda.current_stream().cuda_stream)
        return h

    def cusparseDestroy(self, h):
        self._call(self.real, "cusparseDestroy", h)

    def cusparseSetStream(self, h, stream):
        self._call(self.real, "cusparseSetStream", h, C.c_void_p(int(stream)))

    def cusparseSetPointerMode(self, h, mode):
        self._call(self.real, "cusparseSetPointerMode", h, C.c_int(mode))

    # ---- descriptors -------------------------------------------------------------------------------
    def cusparseCreateCsr(self, rows, cols, nnz, csrRowOffsets, csrColInd, csrValues, idxBase=CUSPARSE_INDEX_BASE_ZERO):
        d = C.c_void_p()
        self._call(self.lib, "cusparseCreateCsr", C.byref(d), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz),
                   _ptr(csrRowOffsets), _ptr(csrColInd), _ptr(csrValues), C.c_int(_IT[csrRowOffsets.dtype]),
                   C.c_int(_IT[csrColInd.dtype]), C.c_int(idxBase), C.c_int(_VT[csrValues.dtype]))
        return d

    def cusparseCreateCoo(self, rows, cols, nnz, cooRowInd, cooColInd, cooValues, idxBase=CUSPARSE_INDEX_BASE_ZERO):
        d = C.c_void_p()
        self._call(self.lib, "cusparseCreateCoo", C.byref(d), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz),
                   _ptr(cooRowInd), _ptr(cooColInd), _ptr(cooValues), C.c_int(_IT[cooRowInd.dtype]), C.c_int(idxBase),
                   C.c_int(_VT[cooValues.dtype]))
        return d

    def cusparseCreateSlicedEll(self, rows, cols, nnz, sellValuesSize, sliceSize, sellSliceOffsets, sellColInd, sellValues,
                                idxBase=CUSPARSE_INDEX_BASE_ZERO):
        d = C.c_void_p()
        self._call(self.lib, "cusparseCreateSlicedEll", C.byref(d), C.c_int64(rows), C.c_int64(cols), C.c_int64(nnz),
                   C.c_int64(sellValuesSize), C.c_int64(sliceSize), _ptr(sellSliceOffsets), _ptr(sellColInd),
                   _ptr(sellValues), C.c_int(_IT[sellSliceOffsets.dtype]), C.c_int(_IT[sellColInd.dtype]),
                   C.c_int(idxBase), C.c_int(_VT[sellValues.dtype]))
        return d

    def cusparseDestroySpMat(self, d):
        self._call(self.lib, "cusparseDestroySpMat", d)

    def cusparseCreateDnVec(self, size, values):
        d = C.c_void_p()
        self._call(self.lib, "cusparseCreateDnVec", C.byref(d), C.c_int64(size), _ptr(values), C.c_int(_VT[values.dtype]))
        return d

    def cusparseDestroyDnVec(self, d):
        self._call(self.lib, "cusparseDestroyDnVec", d)

    def cusparseDnVecSetValues(self, d, values):
        self._call(self.lib, "cusparseDnVecSetValues", d, _ptr(values))

    # dense matrices: the descriptors always belong to the real library (the shim reads them through cusparseConstDnMatGet)
    def cusparseCreateDnMat(self, rows, cols, ld, values, order=CUSPARSE_ORDER_COL):
        d = C.c_void_p()
        self._call(self.real, "cusparseCreateDnMat", C.byref(d), C.c_int64(rows), C.c_int64(cols), C.c_int64(ld), _ptr(values),
                   C.c_int(_VT[values.dtype]), C.c_int(order))
        return d

    def cusparseDnMatSetStridedBatch(self, d, batchCount, batchStride):
        self._call(self.real, "cusparseDnMatSetStridedBatch", d, C.c_int(batchCount), C.c_int64(batchStride))

    def cusparseCsrSetStridedBatch(self, d, batchCount, offsetsBatchStride, columnsValuesBatchStride):
        """spmm_csr_batched_example.c:140 (strides in elements; 0 = the array is shared by the whole batch)."""
        self._call(self.lib, "cusparseCsrSetStridedBatch", d, C.c_int(batchCount), C.c_int64(offsetsBatchStride),
                   C.c_int64(columnsValuesBatchStride))

    def cusparseDestroyDnMat(self, d):
        self._call(self.real, "cusparseDestroyDnMat", d)

    def cusparseSpMM_bufferSize(self, handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg=CUSPARSE_SPMM_ALG_DEFAULT):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        size = C.c_size_t(0)
        self._call(self.lib, "cusparseSpMM_bufferSize", handle, C.c_int(opA), C.c_int(opB), pa, matA, matB, pb, matC,
                   C.c_int(computeType), C.c_int(alg), C.byref(size))
        return int(size.value)

    def cusparseSpMM_preprocess(self, handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg, externalBuffer):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        self._call(self.lib, "cusparseSpMM_preprocess", handle, C.c_int(opA), C.c_int(opB), pa, matA, matB, pb, matC,
                   C.c_int(computeType), C.c_int(alg), _ptr(externalBuffer))

    def cusparseSpMM(self, handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg, externalBuffer):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        self._call(self.lib, "cusparseSpMM", handle, C.c_int(opA), C.c_int(opB), pa, matA, matB, pb, matC,
                   C.c_int(computeType), C.c_int(alg), _ptr(externalBuffer))

    def cusparseCsrSetPointers(self, d, off, col, val):
        self._call(self.lib, "cusparseCsrSetPointers", d, _ptr(off), _ptr(col), _ptr(val))

    # ---- SpMV --------------------------------------------------------------------------------------
    @staticmethod
    def _scalar(v, dtype):
        """alpha/beta: python number -> host scalar of the compute type; torch tensor -> device pointer."""
        if isinstance(v, torch.Tensor):
            return C.c_void_p(v.data_ptr()), v
        c = _CT[dtype](v)
        return C.cast(C.pointer(c), C.c_void_p), c

    def cusparseSpMV_bufferSize(self, handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg=CUSPARSE_SPMV_ALG_DEFAULT):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        size = C.c_size_t(0)
        self._call(self.lib, "cusparseSpMV_bufferSize", handle, C.c_int(opA), pa, matA, vecX, pb, vecY, C.c_int(computeType),
                   C.c_int(alg), C.byref(size))
        return int(size.value)

    def cusparseSpMV_preprocess(self, handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        self._call(self.lib, "cusparseSpMV_preprocess", handle, C.c_int(opA), pa, matA, vecX, pb, vecY, C.c_int(computeType),
                   C.c_int(alg), _ptr(externalBuffer))

    def cusparseSpMV(self, handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer):
        dt = torch.float32 if computeType == CUDA_R_32F else torch.float64
        pa, _ka = self._scalar(alpha, dt)
        pb, _kb = self._scalar(beta, dt)
        self._call(self.lib, "cusparseSpMV", handle, C.c_int(opA), pa, matA, vecX, pb, vecY, C.c_int(computeType),
                   C.c_int(alg), _ptr(externalBuffer))


class SpMVOperator:
    """The sample's call sequence packaged once: descriptors + external buffer live as long as the operator
    (cg_example.c:387-418 keeps matA / d_bufferMV for the whole solve and calls cusparseSpMV per iteration)."""

    def __init__(self, api: Api, fmt: str, rows: int, cols: int, arrays: dict, base: int = 0, preprocess: bool = True,
                 alg: int = CUSPARSE_SPMV_ALG_DEFAULT, handle=None, op: int = CUSPARSE_OPERATION_NON_TRANSPOSE,
                 xy_dtype: torch.dtype | None = None):
        """xy_dtype: type of x, y, alpha, beta and of the arithmetic when it differs from the type of A's values (mixed
        precision: fp32 A with fp64 vectors); index widths follow the dtypes of the index tensors (int32 / int64)."""
        self.api, self.fmt, self.rows, self.cols, self.base, self.alg, self.op = api, fmt, rows, cols, base, alg, op
        self.arrays = arrays  # keeps the device tensors alive
        self.own_handle = handle is None
        self.handle = api.cusparseCreate() if handle is None else handle
        val = arrays["val"]
        self.dtype = xy_dtype or val.dtype      # the type of x / y / the scalars / the arithmetic
        self.ctype = _VT[self.dtype]
        if fmt == "csr":
            self.nnz = int(arrays["col"].numel())
            self.mat = api.cusparseCreateCsr(rows, cols, self.nnz, arrays["off"], arrays["col"], val, base)
        elif fmt == "coo":
            self.nnz = int(arrays["col"].numel())
            self.mat = api.cusparseCreateCoo(rows, cols, self.nnz, arrays["row"], arrays["col"], val, base)
        elif fmt == "sell":
            self.nnz = int(arrays["nnz"])
            self.mat = api.cusparseCreateSlicedEll(rows, cols, self.nnz, int(val.numel()), int(arrays["slice_size"]),
                                                   arrays["off"], arrays["col"], val, base)
        else:
            raise ValueError(fmt)
        nx, ny = (cols, rows) if op == CUSPARSE_OPERATION_NON_TRANSPOSE else (rows, cols)      # A^T: y[cols] = A^T x[rows]
        self._x = torch.empty(max(nx, 1), dtype=self.dtype, device=val.device)
        self._y = torch.empty(max(ny, 1), dtype=self.dtype, device=val.device)
        self.vecX = api.cusparseCreateDnVec(nx, self._x)
        self.vecY = api.cusparseCreateDnVec(ny, self._y)
        size = api.cusparseSpMV_bufferSize(self.handle, op, 1.0, self.mat, self.vecX, 0.0,
                                           self.vecY, self.ctype, alg)
        self.buffer_bytes = size
        self.buffer = torch.empty(max(size, 16), dtype=torch.uint8, device=val.device)
        if preprocess:
            api.cusparseSpMV_preprocess(self.handle, op, 1.0, self.mat, self.vecX, 0.0,
                                        self.vecY, self.ctype, alg, self.buffer)

    def __call__(self, x: torch.Tensor, y: torch.Tensor, alpha=1.0, beta=0.0):
        """y = alpha*A*x + beta*y, asynchronous on the handle's stream."""
        a = self.api
        a.cusparseDnVecSetValues(self.vecX, x)
        a.cusparseDnVecSetValues(self.vecY, y)
        a.cusparseSpMV(self.handle, self.op, alpha, self.mat, self.vecX, beta, self.vecY, self.ctype,
                       self.alg, self.buffer)
        return y

    def prebuilt(self, x: torch.Tensor, y: torch.Tensor, alpha=1.0, beta=0.0):
        """A zero-argument callable that issues exactly this cusparseSpMV (fixed x, y, alpha, beta) with the ctypes
        arguments built once: ~2 us of host time per call instead of ~15 (timing loops, CUDA-graph capture)."""
        a = self.api
        a.cusparseDnVecSetValues(self.vecX, x)
        a.cusparseDnVecSetValues(self.vecY, y)
        ct = _CT[self.dtype]
        ca, cb = ct(alpha), ct(beta)
        fn = a.lib.cusparseSpMV
        argv = (self.handle, C.c_int(self.op), C.cast(C.pointer(ca), C.c_void_p), self.mat, self.vecX,
                C.cast(C.pointer(cb), C.c_void_p), self.vecY, C.c_int(self.ctype), C.c_int(self.alg),
                C.c_void_p(self.buffer.data_ptr()))
        vx, vy, setv = self.vecX, self.vecY, a.lib.cusparseDnVecSetValues
        px, py = C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())

        def call(_keep=(ca, cb, x, y)):
            setv(vx, px)            # several prebuilt calls may share this operator's vector descriptors
            setv(vy, py)
            st = fn(*argv)
            if st != 0:
                raise CuSparseError("cusparseSpMV", st)
        return call

    def close(self):
        if self.mat is not None:
            self.api.cusparseDestroySpMat(self.mat)
            self.api.cusparseDestroyDnVec(self.vecX)
            self.api.cusparseDestroyDnVec(self.vecY)
            if self.own_handle:
                self.api.cusparseDestroy(self.handle)
            self.mat = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmm(api: Api, rows: int, cols: int, arrays: dict, B: torch.Tensor, C0: torch.Tensor, alpha=1.0, beta=0.0,
         order_b: int = CUSPARSE_ORDER_COL, order_c: int = CUSPARSE_ORDER_COL, base: int = 0, timing=None) -> torch.Tensor:
    """The call sequence of cuSPARSE/spmm_csr/spmm_csr_example.c:86-132 once: C = alpha*A*B + beta*C0.

    B and C0 are 1-D device buffers holding the cols x n / rows x n matrices in the given order with the tight leading
    dimension (column-major: ld = rows of the matrix; row-major: ld = n).  Returns the C buffer."""
    val = arrays["val"]
    n = B.numel() // max(cols, 1)
    ct = _VT[val.dtype]
    h = api.cusparseCreate()
    matA = api.cusparseCreateCsr(rows, cols, int(arrays["col"].numel()), arrays["off"], arrays["col"], val, base)
    Cb = C0.clone()
    matB = api.cusparseCreateDnMat(cols, n, cols if order_b == CUSPARSE_ORDER_COL else n, B, order_b)
    matC = api.cusparseCreateDnMat(rows, n, rows if order_c == CUSPARSE_ORDER_COL else n, Cb, order_c)
    op = CUSPARSE_OPERATION_NON_TRANSPOSE
    size = api.cusparseSpMM_bufferSize(h, op, op, alpha, matA, matB, beta, matC, ct)
    buf = torch.empty(max(size, 16), dtype=torch.uint8, device=val.device)
    api.cusparseSpMM_preprocess(h, op, op, alpha, matA, matB, beta, matC, ct, CUSPARSE_SPMM_ALG_DEFAULT, buf)
    if timing is not None:          # (start, stop) CUDA events around the cusparseSpMM call alone (scripts/bench_formats.py)
        timing[0].record()
    api.cusparseSpMM(h, op, op, alpha, matA, matB, beta, matC, ct, CUSPARSE_SPMM_ALG_DEFAULT, buf)
    if timing is not None:
        timing[1].record()
    torch.cuda.synchronize()
    api.cusparseDestroySpMat(matA)
    api.cusparseDestroyDnMat(matB)
    api.cusparseDestroyDnMat(matC)
    api.cusparseDestroy(h)
    return Cb


def spmm_batched(api: Api, rows: int, cols: int, nnz: int, batches: int, off: torch.Tensor, col: torch.Tensor, val: torch.Tensor,
                 B: torch.Tensor, C0: torch.Tensor, alpha=1.0, beta=0.0, off_stride: int = 0, colval_stride: int | None = None,
                 b_stride: int | None = None, order: int = CUSPARSE_ORDER_COL) -> torch.Tensor:
    """The call sequence of cuSPARSE/spmm_csr_batched/spmm_csr_batched_example.c:128-160: C_i = alpha*A_i*B_i + beta*C_i.

    off / col / val hold the batch back to back with the given element strides (off_stride 0: shared row offsets, as in the
    sample; colval_stride 0: the whole matrix shared, the sample's "matA broadcast" variant); B and C0 are 1-D buffers of
    `batches` dense matrices with the tight leading dimension (b_stride 0: B shared)."""
    n = C0.numel() // (batches * max(rows, 1))
    ct = _VT[val.dtype]
    colval_stride = nnz if colval_stride is None else colval_stride
    b_stride = cols * n if b_stride is None else b_stride
    h = api.cusparseCreate()
    matA = api.cusparseCreateCsr(rows, cols, nnz, off, col, val)
    api.cusparseCsrSetStridedBatch(matA, batches, off_stride, colval_stride)
    Cb = C0.clone()
    matB = api.cusparseCreateDnMat(cols, n, cols if order == CUSPARSE_ORDER_COL else n, B, order)
    if b_stride:
        api.cusparseDnMatSetStridedBatch(matB, batches, b_stride)
    matC = api.cusparseCreateDnMat(rows, n, rows if order == CUSPARSE_ORDER_COL else n, Cb, order)
    api.cusparseDnMatSetStridedBatch(matC, batches, rows * n)
    op = CUSPARSE_OPERATION_NON_TRANSPOSE
    alg = CUSPARSE_SPMM_CSR_ALG2                        # as in the sample (:150,157)
    size = api.cusparseSpMM_bufferSize(h, op, op, alpha, matA, matB, beta, matC, ct, alg)
    buf = torch.empty(max(size, 16), dtype=torch.uint8, device=val.device)
    api.cusparseSpMM(h, op, op, alpha, matA, matB, beta, matC, ct, alg, buf)
    torch.cuda.synchronize()
    api.cusparseDestroySpMat(matA)
    api.cusparseDestroyDnMat(matB)
    api.cusparseDestroyDnMat(matC)
    api.cusparseDestroy(h)
    return Cb
