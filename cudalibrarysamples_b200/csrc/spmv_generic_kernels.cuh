// spmv_generic_kernels.cuh -- the device code of spmv_generic.cu (see the comment there), kept in a header of its own so that
// tests/host_emulation/ can compile THE SAME SOURCE for the host (-DB200_HOST_EMULATION: threads stand in for lanes, a barrier
// based __shfl_down_sync, a locked atomicAdd) and run every kernel against the CPU oracle without a GPU
// (tests/test_generic_emulation.py).  Nothing here may depend on anything but spmv_common.cuh's Scalars / axpby.
#pragma once
#ifdef B200_HOST_EMULATION
#include "cuda_emulation.h"            // tests/host_emulation/: threadIdx / blockIdx / gridDim, __shfl_down_sync, atomicAdd, Scalars, axpby
#else
#include "spmv_common.cuh"
#endif

namespace b200 {

constexpr int GEN_BLOCK = 256;
constexpr int GEN_MAX_CTAS = 148 * 16;           // grid-stride kernels: a few waves of the 148 SMs

// OffT: CSR row offsets / Sliced-ELL slice offsets / COO row indices; ColT: column indices; AT: values of A; XT: x, y,
// alpha, beta and the arithmetic.
template <typename OffT, typename ColT, typename AT, typename XT>
struct GenArgs {
    const OffT* off;
    const ColT* col;
    const AT*   val;
    const XT*   x;
    XT*         y;
    long long   rows, cols, nnz, base, slice_size;
    Scalars<XT> s;
};

template <typename T>
__global__ void __launch_bounds__(GEN_BLOCK) gen_scale_y_kernel(T* __restrict__ y, long long n, Scalars<T> s) {
    const T beta = s.b();
    if (beta == T(1)) return;
    for (long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * GEN_BLOCK)
        y[i] = beta == T(0) ? T(0) : beta * y[i];                  // beta == 0 never reads y
}

template <typename OffT, typename ColT, typename AT, typename XT>
__global__ void __launch_bounds__(GEN_BLOCK) csr_generic_kernel(const GenArgs<OffT, ColT, AT, XT> a, int lanes_log2) {
    const int       lanes = 1 << lanes_log2;
    const int       lane = (int)threadIdx.x & 31;
    const int       sub = lane & (lanes - 1);
    const long long rows_per_warp = 32 >> lanes_log2;
    const long long warp = ((long long)blockIdx.x * GEN_BLOCK + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * GEN_BLOCK) >> 5;
    const XT        alpha = a.s.a(), beta = a.s.b();
    // w0 depends on the warp only: all 32 lanes make the same number of trips, so the full-mask shuffles below are legal
    for (long long w0 = warp * rows_per_warp; w0 < a.rows; w0 += nwarps * rows_per_warp) {
        const long long row = w0 + (lane >> lanes_log2);
        XT sum = XT(0);
        if (row < a.rows) {
            const long long beg = (long long)a.off[row] - a.base, end = (long long)a.off[row + 1] - a.base;
            for (long long k = beg + sub; k < end; k += lanes) {
                const long long c = (long long)a.col[k] - a.base;
                sum += (XT)a.val[k] * a.x[c];
            }
        }
        for (int o = lanes >> 1; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, lanes);
        if (row < a.rows && sub == 0) {
            XT* yp = a.y + row;
            *yp = axpby(alpha, sum, beta, yp);
        }
    }
}

template <typename OffT, typename ColT, typename AT, typename XT>
__global__ void __launch_bounds__(GEN_BLOCK) csr_generic_transpose_kernel(const GenArgs<OffT, ColT, AT, XT> a, int lanes_log2) {
    const int       lanes = 1 << lanes_log2;
    const int       sub = (int)threadIdx.x & (lanes - 1);
    const long long group = ((long long)blockIdx.x * GEN_BLOCK + threadIdx.x) >> lanes_log2;
    const long long ngroups = ((long long)gridDim.x * GEN_BLOCK) >> lanes_log2;
    const XT        alpha = a.s.a();
    for (long long row = group; row < a.rows; row += ngroups) {
        const long long beg = (long long)a.off[row] - a.base, end = (long long)a.off[row + 1] - a.base;
        if (end <= beg) continue;
        const XT xr = a.x[row];
        for (long long k = beg + sub; k < end; k += lanes)
            atomicAdd(a.y + ((long long)a.col[k] - a.base), alpha * (XT)a.val[k] * xr);
    }
}

// COO: a.off = row indices.  (The caller swaps the index arrays and rows / cols for A^T.)
template <typename OffT, typename ColT, typename AT, typename XT>
__global__ void __launch_bounds__(GEN_BLOCK) coo_generic_kernel(const GenArgs<OffT, ColT, AT, XT> a) {
    const XT alpha = a.s.a();
    for (long long i = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x; i < a.nnz; i += (long long)gridDim.x * GEN_BLOCK) {
        const long long r = (long long)a.off[i] - a.base, c = (long long)a.col[i] - a.base;
        atomicAdd(a.y + r, alpha * (XT)a.val[i] * a.x[c]);
    }
}

// Sliced-ELL (spmv_sell_example.c:48-66): slice s of width w holds element (row r of the slice, k) at
// sliceOff[s] + k * sliceSize + r; padding entries carry column -1 (+ base).
template <typename OffT, typename ColT, typename AT, typename XT, bool TRANSPOSE>
__global__ void __launch_bounds__(GEN_BLOCK) sell_generic_kernel(const GenArgs<OffT, ColT, AT, XT> a) {
    const XT        alpha = a.s.a(), beta = a.s.b();
    const long long S = a.slice_size;
    for (long long row = (long long)blockIdx.x * GEN_BLOCK + threadIdx.x; row < a.rows; row += (long long)gridDim.x * GEN_BLOCK) {
        const long long s = row / S, r = row - s * S;
        const long long beg = (long long)a.off[s] - a.base, end = (long long)a.off[s + 1] - a.base;
        const long long width = (end - beg) / S;
        const XT xr = TRANSPOSE ? a.x[row] : XT(0);
        XT sum = XT(0);
        for (long long k = 0; k < width; k++) {
            const long long i = beg + k * S + r;
            const long long c = (long long)a.col[i] - a.base;
            if (c < 0) continue;                                    // padding
            if (TRANSPOSE) atomicAdd(a.y + c, alpha * (XT)a.val[i] * xr);
            else sum += (XT)a.val[i] * a.x[c];
        }
        if (!TRANSPOSE) {
            XT* yp = a.y + row;
            *yp = axpby(alpha, sum, beta, yp);
        }
    }
}

}  // namespace b200
