"""CPU: pin the oracle (oracle/spmv_oracle.c) against every golden the reference holds for the SpMV path."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O

T = O.TOY


def test_toy_csr_golden():
    # cuSPARSE/spmv_csr/spmv_csr_example.c:45-56,123-129: exact equality with {19, 8, 51, 52}
    y = O.spmv_csr(T["csr_off"], T["csr_col"], T["val"], T["x"], alpha=1.0, beta=0.0)
    assert y.dtype == np.float32 and np.array_equal(y, T["y_result"])


def test_toy_coo_golden():
    # cuSPARSE/spmv_coo/spmv_coo_example.c:48-54,115-121
    y = O.spmv_coo(4, T["coo_row"], T["csr_col"], T["val"], T["x"])
    assert np.array_equal(y, T["y_result"])


def test_toy_sell_golden():
    # cuSPARSE/spmv_sell/spmv_sell_example.c:48-69,133-139 (slice size 2, padding column -1)
    y = O.spmv_sell(4, 2, T["sell_off"], T["sell_col"], T["sell_val"], T["x"])
    assert np.array_equal(y, T["y_result"])


def test_spmvop_alpha_beta_golden():
    # cuSPARSE/spmvop_csr/spmv_csr_op_example.c:169-173,288-289,307-318: fp64, alpha=1, beta=3, y={5,6,7,8};
    # the host loop of the sample gives {19+15, 8+18, 51+21, 52+24}.
    y = O.spmv_csr(T["csr_off"], T["csr_col"], T["val"].astype(np.float64), T["x"].astype(np.float64),
                   y=np.array([5.0, 6.0, 7.0, 8.0]), alpha=1.0, beta=3.0)
    assert np.array_equal(y, np.array([34.0, 26.0, 72.0, 76.0]))


def test_converters_reproduce_the_samples_layouts():
    # CSR of spmv_csr_example.c -> the COO / SELL arrays hard-coded in spmv_coo_example.c / spmv_sell_example.c
    assert np.array_equal(O.csr_to_coo_rows(T["csr_off"]), T["coo_row"])
    so, co, vo = O.csr_to_sell(T["csr_off"], T["csr_col"], T["val"], 2)
    assert np.array_equal(so, T["sell_off"]) and np.array_equal(co, T["sell_col"]) and np.array_equal(vo, T["sell_val"])


def test_cg_generator_matches_readme():
    # cuSPARSE/cg/README.md:72-75 prints rows 490000, nnz 2447200 for grid 700 (cg_example.c:75,82)
    off, col, val = O.gen_stencil5(700)
    assert off.size - 1 == 490000 and col.size == 2447200 and off[-1] == 2447200
    # interior row: -1 -1 4.04 -1 -1, sorted columns (cg_example.c:100,118-122)
    r = 700 * 3 + 5
    assert np.array_equal(col[off[r]:off[r + 1]], [r - 700, r - 1, r, r + 1, r + 700])
    assert np.allclose(val[off[r]:off[r + 1]], [-1, -1, 4.04, -1, -1])
    A = sp.csr_matrix((val, col, off), shape=(490000, 490000))
    assert abs(A - A.T).max() == 0.0


def test_bicgstab_generator():
    # cuSPARSE/bicgstab/bicgstab_example.c:96-98,117-121: mass .3, ux .3, uy .2
    off, col, val = O.gen_stencil5(50, mass=0.3, ux=0.3, uy=0.2)
    assert col.size == 5 * 2500 - 4 * 50
    r = 50 * 7 + 9
    assert np.allclose(val[off[r]:off[r + 1]], [-1.3, -1.2, 4.8, -1.0, -1.0])


def test_laplace7_generator():
    # cuDSS/simple_residual/laplace_generator.hxx:34-107: nnz = 7 n^3 - 6 n^2, diag 16
    nx = 12
    off, col, val = O.gen_laplace7(nx)
    assert col.size == 7 * nx ** 3 - 6 * nx ** 2
    A = sp.csr_matrix((val, col, off), shape=(nx ** 3,) * 2)
    assert np.all(A.diagonal() == 16.0) and abs(A - A.T).max() == 0.0
    assert np.all(np.diff(col[off[100]:off[101]]) > 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (0.75, 0.0), (2.5, -0.5)])
def test_csr_against_scipy(dtype, alpha, beta):
    # (alpha, beta) pairs are the ones the solvers use: cg_example.c:153-160,220-224,405-418
    off, col, val = O.rmat_csr(3000, avg_nnz=8, seed=1, val_seed=2, dtype=dtype)
    x = O.uniform(3, 3000, dtype)
    y0 = O.uniform(4, 3000, dtype)
    A = sp.csr_matrix((val.astype(np.float64), col, off), shape=(3000, 3000))
    want = alpha * (A @ x.astype(np.float64)) + beta * y0.astype(np.float64)
    got = O.spmv_csr(off, col, val, x, y0, alpha, beta)
    tol = 1e-13 if dtype == np.float64 else 1e-6
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol
    got_mt = O.spmv_csr(off, col, val, x, y0, alpha, beta, threads=4)
    assert np.array_equal(got, got_mt)


def test_formats_agree_and_base_one():
    off, col, val = O.rmat_csr(2000, avg_nnz=6, seed=5, val_seed=6)
    x = O.uniform(7, 2000)
    y = O.spmv_csr(off, col, val, x)
    assert np.array_equal(y, O.spmv_csr(off + 1, col + 1, val, x, base=1))
    row = O.csr_to_coo_rows(off)
    assert np.allclose(y, O.spmv_coo(2000, row, col, val, x), rtol=0, atol=1e-13)
    for ss in (1, 2, 32, 7):
        so, sc, sv = O.csr_to_sell(off, col, val, ss)
        assert np.array_equal(y, O.spmv_sell(2000, ss, so, sc, sv, x))


def test_rmat_is_deterministic_and_well_formed():
    a = O.rmat_csr(5000, avg_nnz=16, seed=42, val_seed=43)
    b = O.rmat_csr(5000, avg_nnz=16, seed=42, val_seed=43)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    off, col, val = a
    assert off[0] == 0 and off[-1] == col.size == val.size
    assert col.size <= 5000 * 16 and col.min() >= 0 and col.max() < 5000
    # columns strictly increasing inside each row (duplicates merged, sorted)
    d = np.diff(col.astype(np.int64))
    row_start = np.zeros(col.size, bool)
    row_start[off[:-1][off[:-1] < col.size]] = True
    assert np.all(d[~row_start[1:]] > 0)
    assert np.all(np.abs(val) <= 1.0)
    # skew: R-MAT leaves many rows empty and a few very long
    lens = np.diff(off)
    assert (lens == 0).mean() > 0.2 and lens.max() > 20 * lens.mean()


def test_empty_and_ragged():
    off = np.array([0, 0, 0, 2, 2, 5, 5], np.int32)
    col = np.array([1, 5, 0, 2, 3], np.int32)
    val = np.arange(1, 6, dtype=np.float64)
    x = np.arange(1, 7, dtype=np.float64)
    y0 = np.full(6, 10.0)
    y = O.spmv_csr(off, col, val, x, y0, 2.0, 0.5)
    assert np.array_equal(y, [5, 5, 2 * (2 + 12) + 5, 5, 2 * (3 + 12 + 20) + 5, 5])
    # zero-size
    assert O.spmv_csr(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(0)).size == 0


def test_spmm_golden_of_the_reference_sample():
    """cuSPARSE/spmm_csr/spmm_csr_example.c:59-66: hB, hC_result (column-major 4x3), exact `!=` compare at :143-151."""
    T = O.TOY
    want = np.array([19, 8, 51, 52, 43, 24, 123, 120, 67, 40, 195, 188], np.float32)
    for ob in ("col", "row"):
        for oc in ("col", "row"):
            C = O.spmm_csr(T["csr_off"], T["csr_col"], T["val"], T["spmm_B"], order_b=ob, order_c=oc)
            assert np.array_equal(np.asfortranarray(C).T.reshape(-1), want), (ob, oc)
    # alpha / beta and fp64, against scipy
    import scipy.sparse as sp
    off, col, val = O.rmat_csr(500, avg_nnz=6, seed=1, val_seed=2)
    A = sp.csr_matrix((val, col, off), shape=(500, 500))
    rng = np.random.default_rng(0)
    B, C0 = rng.uniform(-1, 1, (500, 7)), rng.uniform(-1, 1, (500, 7))
    got = O.spmm_csr(off, col, val, B, C0, alpha=-0.5, beta=2.0, order_b="row", order_c="col", threads=2)
    assert np.allclose(got, -0.5 * (A @ B) + 2.0 * C0, rtol=1e-13, atol=1e-13)


def test_spmm_batched_golden_of_the_reference_sample():
    """cuSPARSE/spmm_csr_batched/spmm_csr_batched_example.c:56-88: two products sharing the row offsets; hC1_result / hC2_result."""
    T = O.TOY_BATCHED
    for i in range(T["batches"]):
        B = T["B"][i].reshape(T["n"], T["cols"]).T              # the column-major 4x3 buffer as a 2-D array
        C = O.spmm_csr(T["csr_off"], T["csr_col"][i], T["val"][i], B)
        assert np.array_equal(np.asfortranarray(C).T.reshape(-1), T["C"][i]), i
