"""CPU: the shim's descriptor side (no device needed -- the real library's descriptor calls are host-only): descriptors created
through the shim's cusparseCreateCsr stay real descriptors, cusparseCsrSetStridedBatch reaches the real library AND the side
table, and the strided-batch decision of cusparseSpMM (b200spmm_batch_count) follows B200SPMV_GENERIC.  Runs in a subprocess:
the shim picks its real libcusparse once per process (B200SPMV_CUSPARSE)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REAL = "/usr/local/cuda/lib64/libcusparse.so.12"

SCRIPT = textwrap.dedent("""
    import ctypes as C, sys
    real = C.CDLL(sys.argv[2], mode=C.RTLD_GLOBAL)
    shim = C.CDLL(sys.argv[1])
    I32, F32, COL = 2, 0, 1
    off = (C.c_int * 5)(0, 3, 4, 7, 9)
    col = (C.c_int * 18)()
    val = (C.c_float * 18)()
    B = (C.c_float * 24)()
    Cm = (C.c_float * 24)()

    def csr(lib):
        d = C.c_void_p()
        assert lib.cusparseCreateCsr(C.byref(d), C.c_int64(4), C.c_int64(4), C.c_int64(9), off, col, val, I32, I32, 0, F32) == 0
        return d

    def dn(batches, stride):
        d = C.c_void_p()
        assert real.cusparseCreateDnMat(C.byref(d), C.c_int64(4), C.c_int64(3), C.c_int64(4), B, F32, COL) == 0
        if batches > 1:
            assert real.cusparseDnMatSetStridedBatch(d, batches, C.c_int64(stride)) == 0
        return d

    def count(a, b, c):
        return shim.b200spmm_batch_count(a, b, c)

    def mode(m):
        assert shim.b200spmv_set_option(b"B200SPMV_GENERIC", m) == 0

    A = csr(shim)
    n = C.c_int(-1)
    assert real.cusparseSpMatGetStridedBatch(A, C.byref(n)) == 0 and n.value == 1     # a real descriptor, not batched
    for m in (b"off", b"csr", b"all"):
        mode(m)
        assert count(A, dn(1, 0), dn(1, 0)) == 1                                       # the ordinary product, whatever the switch says
    # spmm_csr_batched_example.c:140-148: shared offsets, per-batch columns / values, B and C strided
    assert shim.cusparseCsrSetStridedBatch(A, 2, C.c_int64(0), C.c_int64(9)) == 0
    assert real.cusparseSpMatGetStridedBatch(A, C.byref(n)) == 0 and n.value == 2     # the real descriptor saw it too
    mode(b"csr")
    assert count(A, dn(2, 12), dn(2, 12)) == 0                                         # default: batches go to the closed library
    mode(b"all")
    assert count(A, dn(2, 12), dn(2, 12)) == 2
    assert count(A, dn(1, 0), dn(2, 12)) == 2                                          # B shared by the batch
    assert count(A, dn(2, 12), dn(3, 12)) == 0                                         # 2 matrices, 3 outputs
    assert count(A, dn(3, 12), dn(2, 12)) == 0
    d = C.c_void_p()                                                                   # overlapping outputs: the real library refuses the stride itself
    assert real.cusparseCreateDnMat(C.byref(d), C.c_int64(4), C.c_int64(3), C.c_int64(4), B, F32, COL) == 0
    assert real.cusparseDnMatSetStridedBatch(d, 2, C.c_int64(6)) != 0
    A1 = csr(shim)                                                                     # one matrix, batched right-hand sides (":141-142 broadcast")
    assert count(A1, dn(2, 12), dn(2, 12)) == 2
    # a batch set behind the shim's back (the real symbol called directly): strides unknown -> not ours
    A2 = csr(shim)
    assert real.cusparseCsrSetStridedBatch(A2, 2, C.c_int64(0), C.c_int64(9)) == 0
    assert count(A2, dn(2, 12), dn(2, 12)) == 0
    # a descriptor the shim never saw: found through the real getters, ordinary product only
    A3 = csr(real)
    assert count(A3, dn(1, 0), dn(1, 0)) == 1
    # every descriptor kind the samples create goes through the shim and stays a REAL descriptor: the real getters see what
    # the shim was given (spmv_coo_example.c:86-89, spmv_sell_example.c:103-107, spmv_csr_example.c:93-95), set-pointer calls reach both sides
    fmt = C.c_int(-1)
    coo = C.c_void_p()
    assert shim.cusparseCreateCoo(C.byref(coo), C.c_int64(4), C.c_int64(4), C.c_int64(9), col, col, val, I32, 0, F32) == 0
    assert real.cusparseSpMatGetFormat(coo, C.byref(fmt)) == 0 and fmt.value == 3
    sell = C.c_void_p()
    assert shim.cusparseCreateSlicedEll(C.byref(sell), C.c_int64(4), C.c_int64(4), C.c_int64(9), C.c_int64(12), C.c_int64(2), off, col, val,
                                        I32, I32, 0, F32) == 0
    assert real.cusparseSpMatGetFormat(sell, C.byref(fmt)) == 0 and fmt.value == 7
    vec = C.c_void_p()
    assert shim.cusparseCreateDnVec(C.byref(vec), C.c_int64(4), B, F32) == 0
    size, vals, vt = C.c_int64(), C.c_void_p(), C.c_int()
    assert real.cusparseDnVecGet(vec, C.byref(size), C.byref(vals), C.byref(vt)) == 0 and size.value == 4 and vals.value == C.addressof(B)
    assert shim.cusparseDnVecSetValues(vec, Cm) == 0
    assert real.cusparseDnVecGetValues(vec, C.byref(vals)) == 0 and vals.value == C.addressof(Cm)
    off2 = (C.c_int * 5)(0, 1, 2, 3, 4)
    assert shim.cusparseCsrSetPointers(A1, off2, col, val) == 0
    r, c_, z, po, pc, pv = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    ot, ct, ib, vt2 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert real.cusparseCsrGet(A1, C.byref(r), C.byref(c_), C.byref(z), C.byref(po), C.byref(pc), C.byref(pv), C.byref(ot), C.byref(ct),
                               C.byref(ib), C.byref(vt2)) == 0 and po.value == C.addressof(off2) and (r.value, c_.value, z.value) == (4, 4, 9)
    assert shim.cusparseDestroyDnVec(vec) == 0
    for d in (coo, sell):
        assert shim.cusparseDestroySpMat(d) == 0
    # what cusparseSpMV would do with descriptors like these (b200spmv_route), under the default switch
    mode(b"csr")
    assert shim.b200spmv_route(1, 0, 0, I32, I32, F32, F32, F32, F32, C.c_int64(4), C.c_int64(4), C.c_int64(9)) == 1
    for d in (A, A1, A2):
        assert shim.cusparseDestroySpMat(d) == 0
    assert real.cusparseDestroySpMat(A3) == 0
    print("OK")
""")


def test_descriptor_side_table_and_batch_decision(built_lib):
    import pytest
    if not os.path.exists(REAL):
        pytest.skip("no toolkit libcusparse here")
    env = {k: v for k, v in os.environ.items() if not k.startswith("B200SPMV_")}
    env["B200SPMV_CUSPARSE"] = REAL
    p = subprocess.run([sys.executable, "-c", SCRIPT, built_lib, REAL], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout + p.stderr
