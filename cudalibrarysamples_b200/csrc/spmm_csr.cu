// spmm_csr.cu -- CSR x dense  C = alpha*A*B + beta*C  on B200 (sm_100a), fp32 / fp64, int32 indices (SURVEY.md 8(f)-1,
// BASELINE.json configs[4]).  Replaces cusparse::csrmm kernels behind cusparseSpMM for CSR descriptors with
// opA = opB = NON_TRANSPOSE (call site: cuSPARSE/spmm_csr/spmm_csr_example.c:105-132; golden C at :64-66).
//
// One warp per row of A and per panel of 64 columns of B / C.  The row's (col, val) pairs are loaded 32 at a time with
// one coalesced instruction each and broadcast with shuffles; for every non-zero the 32 lanes read 64 consecutive
// entries of row `col` of B.  With ROW-major B that is one contiguous 256 B (fp32) segment per non-zero -- the fast case
// SURVEY.md names; with COLUMN-major B (the sample's layout) the same code walks B with stride ldb: correct, one sector per
// element, slower.  C is written once, alpha / beta applied in the epilogue (beta == 0 never reads C).  No tensor cores:
// 2 flop per 4-byte B element fetched.  Bytes per product (row-major, fp32): nnz*8 (A) + nnz*n*4 (B rows, mostly L2 hits
// when B fits) + rows*n*4 (C).
#include "spmv_common.cuh"
#include "config.h"
#include "../../include/b200spmv.h"

namespace b200 {

constexpr int SPMM_WARPS = 8;       // rows per CTA
constexpr int SPMM_PANEL = 64;      // columns of B / C per warp pass (2 per lane)

template <typename T>
struct SpmmArgs {
    const int* off;
    const int* col;
    const T*   val;
    const T*   B;
    T*         C;
    int        base, rows, n;
    long long  sbk, sbj;            // B(k, j) = B[k * sbk + j * sbj]
    long long  sci, scj;            // C(i, j) = C[i * sci + j * scj]
    Scalars<T> s;
};

template <typename T>
__global__ void __launch_bounds__(32 * SPMM_WARPS) spmm_csr_kernel(const SpmmArgs<T> a) {
    const int lane = (int)threadIdx.x & 31;
    const int row = blockIdx.x * SPMM_WARPS + ((int)threadIdx.x >> 5);
    if (row >= a.rows) return;                                   // warp-uniform
    const int j0 = blockIdx.y * SPMM_PANEL;
    const int ja = j0 + lane, jb = j0 + 32 + lane;               // my two columns
    const bool la = ja < a.n, lb = jb < a.n;
    const int b = __ldg(a.off + row) - a.base, e = __ldg(a.off + row + 1) - a.base;
    const T* Ba = a.B + (long long)ja * a.sbj;
    const T* Bb = a.B + (long long)jb * a.sbj;
    T acc0 = T(0), acc1 = T(0);
    for (int p = b; p < e; p += 32) {
        const int  i = p + lane;
        const int  c = i < e ? ldg_stream(a.col + i) - a.base : 0;
        const T    v = i < e ? ldg_stream(a.val + i) : T(0);
        const int  cnt = min(32, e - p);
        for (int t = 0; t < cnt; t += 4) {                       // four B rows in flight per lane
            int kk[4];
            T   vv[4], b0[4], b1[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                kk[u] = __shfl_sync(0xffffffffu, c, (t + u) & 31);
                vv[u] = __shfl_sync(0xffffffffu, v, (t + u) & 31);   // lanes beyond cnt carry v = 0, c = 0: harmless
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long ro = (long long)kk[u] * a.sbk;
                b0[u] = la ? __ldg(Ba + ro) : T(0);
                b1[u] = lb ? __ldg(Bb + ro) : T(0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { acc0 += vv[u] * b0[u]; acc1 += vv[u] * b1[u]; }
        }
    }
    const T alpha = a.s.a(), beta = a.s.b();
    if (la) { T* cp = a.C + (long long)row * a.sci + (long long)ja * a.scj; *cp = axpby(alpha, acc0, beta, cp); }
    if (lb) { T* cp = a.C + (long long)row * a.sci + (long long)jb * a.scj; *cp = axpby(alpha, acc1, beta, cp); }
}

template <typename T>
static int launch_spmm(cudaStream_t stream, int64_t rows, int64_t n, const void* off, const void* col, const void* val, int base,
                       const void* alpha, const void* beta, int on_device, const void* B, int64_t ldb, int b_row_major,
                       void* C, int64_t ldc, int c_row_major) {
    SpmmArgs<T> a;
    a.off = (const int*)off; a.col = (const int*)col; a.val = (const T*)val; a.B = (const T*)B; a.C = (T*)C;
    a.base = base; a.rows = (int)rows; a.n = (int)n;
    a.sbk = b_row_major ? ldb : 1; a.sbj = b_row_major ? 1 : ldb;
    a.sci = c_row_major ? ldc : 1; a.scj = c_row_major ? 1 : ldc;
    if (on_device) { a.s.alpha = T(0); a.s.beta = T(0); a.s.alpha_dev = (const T*)alpha; a.s.beta_dev = (const T*)beta; }
    else { a.s.alpha = *(const T*)alpha; a.s.beta = *(const T*)beta; a.s.alpha_dev = nullptr; a.s.beta_dev = nullptr; }
    const dim3 grid((unsigned)((rows + SPMM_WARPS - 1) / SPMM_WARPS), (unsigned)((n + SPMM_PANEL - 1) / SPMM_PANEL));
    spmm_csr_kernel<T><<<grid, 32 * SPMM_WARPS, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace b200

using namespace b200;

extern "C" int b200spmm_csr(void* stream, int dtype, int64_t rows, int64_t cols, int64_t n, int64_t nnz, const void* row_offsets,
                            const void* col_ind, const void* values, int32_t base, const void* alpha, const void* beta,
                            int scalars_on_device, const void* B, int64_t ldb, int b_row_major, void* C, int64_t ldc,
                            int c_row_major) {
    if (rows < 0 || cols < 0 || n < 0 || nnz < 0 || !alpha || !beta) return -1;
    if (rows == 0 || n == 0) return 0;
    if (rows > INT32_MAX - 1 || n > INT32_MAX - 64 || nnz > INT32_MAX - 65536 || (n + SPMM_PANEL - 1) / SPMM_PANEL > 65535) return -1;
    if (!row_offsets || !C || (nnz > 0 && (!col_ind || !values || !B))) return -1;
    if (ldb < (b_row_major ? n : cols) || ldc < (c_row_major ? n : rows)) return -1;
    if (dtype == 0)
        return launch_spmm<float>((cudaStream_t)stream, rows, n, row_offsets, col_ind, values, base, alpha, beta, scalars_on_device,
                                  B, ldb, b_row_major, C, ldc, c_row_major);
    if (dtype == 1)
        return launch_spmm<double>((cudaStream_t)stream, rows, n, row_offsets, col_ind, values, base, alpha, beta, scalars_on_device,
                                   B, ldb, b_row_major, C, ldc, c_row_major);
    return -1;
}
