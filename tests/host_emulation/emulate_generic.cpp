// emulate_generic.cpp -- TEST INFRASTRUCTURE: the kernels of cudalibrarysamples_b200/csrc/spmv_generic_kernels.cuh, THE SAME
// SOURCE, compiled for the host on top of cuda_emulation.h and exported for tests/test_generic_emulation.py (ctypes).
// The launch geometry (block size, grid-stride loops) is the device one; `ctas` lets a test use a grid far smaller than the
// matrix so that every grid-stride loop makes several trips.
#define B200_HOST_EMULATION
#include "spmv_generic_kernels.cuh"

using namespace b200;

namespace {

struct Call {
    int         transpose, lanes_log2;
    unsigned    ctas;
    long long   rows, cols, nnz, base, slice_size;
    const void *off, *col, *val, *alpha, *beta, *x;
    void*       y;
};

template <typename OffT, typename ColT, typename AT, typename XT>
GenArgs<OffT, ColT, AT, XT> args_of(const Call& c) {
    GenArgs<OffT, ColT, AT, XT> a;
    a.off = (const OffT*)c.off; a.col = (const ColT*)c.col; a.val = (const AT*)c.val; a.x = (const XT*)c.x; a.y = (XT*)c.y;
    a.rows = c.rows; a.cols = c.cols; a.nnz = c.nnz; a.base = c.base; a.slice_size = c.slice_size;
    a.s.alpha = *(const XT*)c.alpha; a.s.beta = *(const XT*)c.beta; a.s.alpha_dev = nullptr; a.s.beta_dev = nullptr;
    return a;
}

// the launch sequences of spmv_generic.cu (launch_csr_generic / launch_coo_generic / launch_sell_generic), on the emulator
template <typename OffT, typename ColT, typename AT, typename XT>
int run_csr(const Call& c) {
    auto a = args_of<OffT, ColT, AT, XT>(c);
    if (c.transpose) {
        emu::launch(c.ctas, GEN_BLOCK, [&] { gen_scale_y_kernel<XT>(a.y, c.cols, a.s); });
        emu::launch(c.ctas, GEN_BLOCK, [&] { csr_generic_transpose_kernel<OffT, ColT, AT, XT>(a, c.lanes_log2); });
    } else {
        emu::launch(c.ctas, GEN_BLOCK, [&] { csr_generic_kernel<OffT, ColT, AT, XT>(a, c.lanes_log2); });
    }
    return 0;
}
template <typename OffT, typename ColT, typename AT, typename XT>
int run_coo(const Call& c) {
    auto a = args_of<OffT, ColT, AT, XT>(c);
    emu::launch(c.ctas, GEN_BLOCK, [&] { gen_scale_y_kernel<XT>(a.y, c.rows, a.s); });
    emu::launch(c.ctas, GEN_BLOCK, [&] { coo_generic_kernel<OffT, ColT, AT, XT>(a); });
    return 0;
}
template <typename OffT, typename ColT, typename AT, typename XT>
int run_sell(const Call& c) {
    auto a = args_of<OffT, ColT, AT, XT>(c);
    if (c.transpose) {
        emu::launch(c.ctas, GEN_BLOCK, [&] { gen_scale_y_kernel<XT>(a.y, c.cols, a.s); });
        emu::launch(c.ctas, GEN_BLOCK, [&] { sell_generic_kernel<OffT, ColT, AT, XT, true>(a); });
    } else {
        emu::launch(c.ctas, GEN_BLOCK, [&] { sell_generic_kernel<OffT, ColT, AT, XT, false>(a); });
    }
    return 0;
}

#define DISPATCH(FN, off64, col64, a_dtype, xy_dtype, call)                                        \
    switch (((off64) ? 8 : 0) | ((col64) ? 4 : 0) | ((a_dtype) ? 2 : 0) | ((xy_dtype) ? 1 : 0)) {  \
        case 0:  return FN<int32_t, int32_t, float, float>(call);                                  \
        case 1:  return FN<int32_t, int32_t, float, double>(call);                                 \
        case 3:  return FN<int32_t, int32_t, double, double>(call);                                \
        case 8:  return FN<int64_t, int32_t, float, float>(call);                                  \
        case 9:  return FN<int64_t, int32_t, float, double>(call);                                 \
        case 11: return FN<int64_t, int32_t, double, double>(call);                                \
        case 12: return FN<int64_t, int64_t, float, float>(call);                                  \
        case 13: return FN<int64_t, int64_t, float, double>(call);                                 \
        case 15: return FN<int64_t, int64_t, double, double>(call);                                \
        default: return -1;                                                                        \
    }

}  // namespace

extern "C" {

int emu_csr_generic(int off64, int col64, int a_dtype, int xy_dtype, int transpose, int lanes_log2, unsigned ctas, long long rows,
                    long long cols, long long nnz, const void* off, const void* col, const void* val, long long base, const void* alpha,
                    const void* beta, const void* x, void* y) {
    Call c{transpose, lanes_log2, ctas, rows, cols, nnz, base, 0, off, col, val, alpha, beta, x, y};
    DISPATCH(run_csr, off64, col64, a_dtype, xy_dtype, c)
}

int emu_coo_generic(int idx64, int a_dtype, int xy_dtype, unsigned ctas, long long rows, long long cols, long long nnz, const void* row,
                    const void* col, const void* val, long long base, const void* alpha, const void* beta, const void* x, void* y) {
    Call c{0, 0, ctas, rows, cols, nnz, base, 0, row, col, val, alpha, beta, x, y};
    DISPATCH(run_coo, idx64, idx64, a_dtype, xy_dtype, c)
}

int emu_sell_generic(int off64, int col64, int a_dtype, int xy_dtype, int transpose, unsigned ctas, long long rows, long long cols,
                     long long slice_size, const void* off, const void* col, const void* val, long long base, const void* alpha,
                     const void* beta, const void* x, void* y) {
    Call c{transpose, 0, ctas, rows, cols, 0, base, slice_size, off, col, val, alpha, beta, x, y};
    DISPATCH(run_sell, off64, col64, a_dtype, xy_dtype, c)
}

}  // extern "C"
