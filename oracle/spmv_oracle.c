/*
 * oracle/spmv_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic that sits behind the reference's
 * cusparseSpMV call sites (y = alpha*A*x + beta*y for CSR / COO / Sliced-ELL)
 * and of the reference's matrix generators.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load this library; the
 * product (cudalibrarysamples_b200/csrc) never links or calls it.
 *
 * Where each function comes from (paths relative to /root/reference):
 *   csr loop ........ cuSPARSE/spmvop_csr/spmv_csr_op_example.c:307-318
 *                     (the only host-side y=aAx+by loop in the reference:
 *                      sum += alpha*val*x[col] over the row, then sum += beta*y)
 *   coo ordering .... cuSPARSE/spmv_coo/spmv_coo_example.c:48-54 (row-major SoA)
 *   sell layout ..... cuSPARSE/spmv_sell/spmv_sell_example.c:48-69
 *                     (column-major inside a slice, padding col = -1 / val = 0)
 *   5-pt Laplacian .. cuSPARSE/cg/cg_example.c:71-128
 *   5-pt adv-diff ... cuSPARSE/bicgstab/bicgstab_example.c:69-127
 *   7-pt Laplacian .. cuDSS/simple_residual/laplace_generator.hxx:34-107
 *
 * Parity pinning: the arithmetic itself lives in closed libcusparse.so.12
 * (12.5.10.65, CUDA 12.9) which is not in the reference tree.  This restatement
 * is pinned against every golden the reference holds for the path (the 4x4 toy
 * result {19,8,51,52} for CSR/COO/SELL, the spmvop alpha/beta toy, the generator
 * nnz counts printed in cg/README.md and bicgstab/README.md) in
 * tests/test_oracle.py, and against the real cusparseSpMV on the GPU box in
 * tests/test_parity_gpu.py.
 *
 * Accumulation is always in double (for fp32 inputs too) so the oracle is at
 * least as accurate as any summation order a GPU kernel may pick.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

EXPORT int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* CSR: row loop of spmv_csr_op_example.c:307-318                             */
/* ------------------------------------------------------------------------- */
#define DEF_CSR(NAME, T)                                                        \
EXPORT void NAME(int64_t rows, const int32_t* off, const int32_t* col,          \
                 const T* val, int32_t base, double alpha, double beta,         \
                 const T* x, T* y, int threads) {                               \
    if (threads <= 1) {                                                         \
        for (int64_t i = 0; i < rows; i++) {                                    \
            double sum = 0;                                                     \
            for (int64_t j = off[i] - base; j < off[i + 1] - base; ++j) {       \
                int64_t k = col[j] - base;                                      \
                sum += alpha * (double)val[j] * (double)x[k];                   \
            }                                                                   \
            if (beta != 0.0) sum += beta * (double)y[i];                        \
            y[i] = (T)sum;                                                      \
        }                                                                       \
        return;                                                                 \
    }                                                                           \
    _Pragma("omp parallel for schedule(dynamic, 2048) num_threads(threads)")    \
    for (int64_t i = 0; i < rows; i++) {                                        \
        double sum = 0;                                                         \
        for (int64_t j = off[i] - base; j < off[i + 1] - base; ++j) {           \
            int64_t k = col[j] - base;                                          \
            sum += alpha * (double)val[j] * (double)x[k];                       \
        }                                                                       \
        if (beta != 0.0) sum += beta * (double)y[i];                            \
        y[i] = (T)sum;                                                          \
    }                                                                           \
}
DEF_CSR(oracle_spmv_csr_f64, double)
DEF_CSR(oracle_spmv_csr_f32, float)

/* ------------------------------------------------------------------------- */
/* CSR x dense (SpMM): C = alpha*A*B + beta*C.  The reference holds no host loop */
/* for it; this is the CSR row loop above applied per column of B, with the      */
/* sample's layouts (cuSPARSE/spmm_csr/spmm_csr_example.c:50-66: column-major,   */
/* ldb = B rows, ldc = A rows; spmm_csr_op_example.c:195-199 uses row-major).    */
/* Pinned by the sample's golden hC_result (spmm_csr_example.c:64-66).           */
/* ------------------------------------------------------------------------- */
#define DEF_SPMM(NAME, T)                                                       \
EXPORT void NAME(int64_t rows, int64_t n, const int32_t* off, const int32_t* col,\
                 const T* val, int32_t base, double alpha, double beta,         \
                 const T* B, int64_t ldb, int b_row_major, T* C, int64_t ldc,   \
                 int c_row_major, int threads) {                                \
    _Pragma("omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)") \
    for (int64_t i = 0; i < rows; i++) {                                        \
        for (int64_t j = 0; j < n; j++) {                                       \
            double sum = 0;                                                     \
            for (int64_t p = off[i] - base; p < off[i + 1] - base; ++p) {       \
                int64_t k = col[p] - base;                                      \
                double b = (double)(b_row_major ? B[k * ldb + j] : B[k + j * ldb]); \
                sum += alpha * (double)val[p] * b;                              \
            }                                                                   \
            T* c = c_row_major ? &C[i * ldc + j] : &C[i + j * ldc];             \
            if (beta != 0.0) sum += beta * (double)*c;                          \
            *c = (T)sum;                                                        \
        }                                                                       \
    }                                                                           \
}
DEF_SPMM(oracle_spmm_csr_f64, double)
DEF_SPMM(oracle_spmm_csr_f32, float)

/* ------------------------------------------------------------------------- */
/* COO (spmv_coo_example.c:48-54): any order of entries is accepted; the sum   */
/* per row is taken in storage order in double, then alpha/beta applied the    */
/* same way as the CSR loop.                                                   */
/* ------------------------------------------------------------------------- */
#define DEF_COO(NAME, T)                                                        \
EXPORT void NAME(int64_t rows, int64_t nnz, const int32_t* row,                 \
                 const int32_t* col, const T* val, int32_t base, double alpha,  \
                 double beta, const T* x, T* y) {                               \
    double* acc = (double*)calloc((size_t)(rows > 0 ? rows : 1), sizeof(double)); \
    for (int64_t j = 0; j < nnz; j++)                                           \
        acc[row[j] - base] += alpha * (double)val[j] * (double)x[col[j] - base];\
    for (int64_t i = 0; i < rows; i++) {                                        \
        double sum = acc[i];                                                    \
        if (beta != 0.0) sum += beta * (double)y[i];                            \
        y[i] = (T)sum;                                                          \
    }                                                                           \
    free(acc);                                                                  \
}
DEF_COO(oracle_spmv_coo_f64, double)
DEF_COO(oracle_spmv_coo_f32, float)

/* ------------------------------------------------------------------------- */
/* Sliced-ELL (spmv_sell_example.c:48-69): element (r,k) of slice s lives at   */
/* sliceOff[s] + k*sliceSize + (r % sliceSize); padding has column -1 (+base). */
/* ------------------------------------------------------------------------- */
#define DEF_SELL(NAME, T)                                                       \
EXPORT void NAME(int64_t rows, int64_t slice_size, const int32_t* slice_off,    \
                 const int32_t* col, const T* val, int32_t base, double alpha,  \
                 double beta, const T* x, T* y, int threads) {                  \
    _Pragma("omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)") \
    for (int64_t i = 0; i < rows; i++) {                                        \
        int64_t s = i / slice_size, lane = i % slice_size;                      \
        int64_t beg = slice_off[s] - base, end = slice_off[s + 1] - base;       \
        int64_t width = (end - beg) / slice_size;                               \
        double sum = 0;                                                         \
        for (int64_t k = 0; k < width; k++) {                                   \
            int64_t idx = beg + k * slice_size + lane;                          \
            int64_t c = (int64_t)col[idx] - base;                               \
            if (c >= 0) sum += alpha * (double)val[idx] * (double)x[c];         \
        }                                                                       \
        if (beta != 0.0) sum += beta * (double)y[i];                            \
        y[i] = (T)sum;                                                          \
    }                                                                           \
}
DEF_SELL(oracle_spmv_sell_f64, double)
DEF_SELL(oracle_spmv_sell_f32, float)

/* ------------------------------------------------------------------------- */
/* Generators                                                                 */
/* ------------------------------------------------------------------------- */

/* cg_example.c:71-128 (ux=uy=0, mass=0.04 -> values -1,-1,4.04,-1,-1) and
 * bicgstab_example.c:69-127 (mass=0.3, ux=0.3, uy=0.2).  Same insertion order:
 * (i-1,j) (i,j-1) (i,j) (i,j+1) (i+1,j), rows i*grid+j.  Returns nnz
 * (= 5*n - 4*grid).  Pass off==NULL to only count. */
EXPORT int64_t oracle_gen_stencil5(int32_t grid, double mass, double ux, double uy,
                                   int32_t* off, int32_t* col, double* val) {
    int64_t n = (int64_t)grid * grid;
    int64_t nnz = 5 * n - 4 * (int64_t)grid;
    if (!off) return nnz;
    int64_t it = 0, row = 0;
    off[0] = 0;
#define INS(u, v, xv)                                                 \
    if (0 <= (u) && (u) < grid && 0 <= (v) && (v) < grid) {           \
        col[it] = (int32_t)((int64_t)(u) * grid + (v));               \
        val[it] = (xv);                                               \
        ++it;                                                         \
    }
    for (int32_t i = 0; i < grid; ++i)
        for (int32_t j = 0; j < grid; ++j) {
            INS(i - 1, j, -1.0 - ux);
            INS(i, j - 1, -1.0 - uy);
            INS(i, j, 4.0 + mass + ux + uy);
            INS(i, j + 1, -1.0);
            INS(i + 1, j, -1.0);
            off[++row] = (int32_t)it;
        }
#undef INS
    return it;
}

/* laplace_generator.hxx:34-107: 7-point stencil on nx^3, diag 16, off-diag -1,
 * neighbour order z-1, y-1, x-1, diag, x+1, y+1, z+1.  off==NULL -> count. */
EXPORT int64_t oracle_gen_laplace7(int32_t nx, int32_t* off, int32_t* col, double* val) {
    int64_t n = (int64_t)nx * nx * nx;
    int64_t nnz = 0;
    if (!off) {
        for (int32_t z = 0; z < nx; z++)
            for (int32_t y = 0; y < nx; y++)
                for (int32_t x = 0; x < nx; x++) {
                    int c = 7;
                    if (z == 0 || z == nx - 1) c--;
                    if (y == 0 || y == nx - 1) c--;
                    if (x == 0 || x == nx - 1) c--;
                    nnz += c;
                }
        return nnz;
    }
    int64_t it = 0;
    off[0] = 0;
    for (int64_t z = 0; z < nx; z++)
        for (int64_t y = 0; y < nx; y++)
            for (int64_t x = 0; x < nx; x++) {
                int64_t row = (z * nx + y) * nx + x;
                if (z > 0)      { col[it] = (int32_t)(row - (int64_t)nx * nx); val[it++] = -1.0; }
                if (y > 0)      { col[it] = (int32_t)(row - nx);               val[it++] = -1.0; }
                if (x > 0)      { col[it] = (int32_t)(row - 1);                val[it++] = -1.0; }
                col[it] = (int32_t)row; val[it++] = 16.0;
                if (x < nx - 1) { col[it] = (int32_t)(row + 1);                val[it++] = -1.0; }
                if (y < nx - 1) { col[it] = (int32_t)(row + nx);               val[it++] = -1.0; }
                if (z < nx - 1) { col[it] = (int32_t)(row + (int64_t)nx * nx); val[it++] = -1.0; }
                off[row + 1] = (int32_t)it;
            }
    (void)n;
    return it;
}

/* Counter-based hash shared with the device-side workload generator
 * (cudalibrarysamples_b200/csrc/workload_gen.cu) so CPU and GPU produce the
 * same synthetic inputs bit for bit.  splitmix64 finaliser. */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t hash2(uint64_t seed, uint64_t i) { return mix64(mix64(seed) + i); }

/* R-MAT edge stream: edge e, level l draws
 * u = hash2(seed, e*64+l) >> 32 and picks the quadrant by thresholds
 * tA, tAB, tABC (= floor(a*2^32) ...).  Bits are MSB first.  No rejection here:
 * callers drop edges with row>=rows or col>=cols. */
EXPORT void oracle_rmat_edges(uint64_t seed, int64_t e0, int64_t count, int32_t scale,
                              uint64_t tA, uint64_t tAB, uint64_t tABC,
                              int64_t* row_out, int64_t* col_out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < count; k++) {
        uint64_t e = (uint64_t)(e0 + k);
        int64_t r = 0, c = 0;
        for (int32_t l = 0; l < scale; l++) {
            uint64_t u = hash2(seed, e * 64u + (uint64_t)l) >> 32;
            int rb, cb;
            if (u < tA)        { rb = 0; cb = 0; }
            else if (u < tAB)  { rb = 0; cb = 1; }
            else if (u < tABC) { rb = 1; cb = 0; }
            else               { rb = 1; cb = 1; }
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        row_out[k] = r;
        col_out[k] = c;
    }
}

/* U(-1,1) doubles from the hash: ((h>>11) * 2^-53) * 2 - 1. */
EXPORT void oracle_uniform_f64(uint64_t seed, int64_t i0, int64_t count, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < count; k++) {
        uint64_t h = hash2(seed, (uint64_t)(i0 + k));
        out[k] = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
}
EXPORT void oracle_uniform_f32(uint64_t seed, int64_t i0, int64_t count, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < count; k++) {
        uint64_t h = hash2(seed, (uint64_t)(i0 + k));
        out[k] = (float)((double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0);
    }
}

/* ------------------------------------------------------------------------- */
/* Format converters (layout rules from the toy samples)                       */
/* ------------------------------------------------------------------------- */

/* CSR -> COO row indices (spmv_coo_example.c:48-49 is the CSR of
 * spmv_csr_example.c:48-49 with row offsets expanded). */
EXPORT void oracle_csr_to_coo_rows(int64_t rows, const int32_t* off, int32_t base, int32_t* row_out) {
    for (int64_t i = 0; i < rows; i++)
        for (int64_t j = off[i] - base; j < off[i + 1] - base; j++) row_out[j] = (int32_t)i + base;
}

/* CSR -> SELL. Pass col_out==NULL to get slice offsets + values size only.
 * Slice width = longest row in the slice; short rows (and rows past the end of
 * the matrix in the last slice) are padded with col=-1 (+base), val=0
 * (spmv_sell_example.c:52-66). Returns sellValuesSize. */
#define DEF_CSR2SELL(NAME, T)                                                        \
EXPORT int64_t NAME(int64_t rows, const int32_t* off, const int32_t* col,            \
                    const T* val, int32_t base, int64_t slice_size,                  \
                    int32_t* slice_off, int32_t* col_out, T* val_out) {              \
    int64_t nslices = (rows + slice_size - 1) / slice_size;                          \
    int64_t pos = 0;                                                                 \
    for (int64_t s = 0; s < nslices; s++) {                                          \
        int64_t r0 = s * slice_size, r1 = r0 + slice_size;                           \
        if (r1 > rows) r1 = rows;                                                    \
        int64_t w = 0;                                                               \
        for (int64_t r = r0; r < r1; r++) {                                          \
            int64_t len = off[r + 1] - off[r];                                       \
            if (len > w) w = len;                                                    \
        }                                                                            \
        slice_off[s] = (int32_t)(pos + base);                                        \
        if (col_out) {                                                               \
            for (int64_t k = 0; k < w; k++)                                          \
                for (int64_t l = 0; l < slice_size; l++) {                           \
                    int64_t r = r0 + l, idx = pos + k * slice_size + l;              \
                    if (r < rows && k < off[r + 1] - off[r]) {                       \
                        col_out[idx] = col[off[r] - base + k];                       \
                        val_out[idx] = val[off[r] - base + k];                       \
                    } else {                                                         \
                        col_out[idx] = -1 + base;                                    \
                        val_out[idx] = (T)0;                                         \
                    }                                                                \
                }                                                                    \
        }                                                                            \
        pos += w * slice_size;                                                       \
    }                                                                                \
    slice_off[nslices] = (int32_t)(pos + base);                                      \
    return pos;                                                                      \
}
DEF_CSR2SELL(oracle_csr_to_sell_f64, double)
DEF_CSR2SELL(oracle_csr_to_sell_f32, float)

/* ------------------------------------------------------------------------- */
/* Timing helper for bench.py's cpu_baseline / --impl reference leg: runs the  */
/* CSR loop `reps` times and returns the minimum wall seconds.                 */
/* ------------------------------------------------------------------------- */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
EXPORT double oracle_time_csr_f64(int64_t rows, const int32_t* off, const int32_t* col,
                                  const double* val, double alpha, double beta,
                                  const double* x, double* y, int threads, int reps,
                                  double* times_out) {
    double best = 1e300;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        oracle_spmv_csr_f64(rows, off, col, val, 0, alpha, beta, x, y, threads);
        double t = now_s() - t0;
        if (times_out) times_out[r] = t;
        if (t < best) best = t;
    }
    return best;
}
