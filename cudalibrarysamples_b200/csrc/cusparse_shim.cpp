// cusparse_shim.cpp -- the drop-in boundary: re-exports the cuSPARSE generic-API symbols that the reference's
// cuSPARSE/spmv_csr, spmv_coo, spmv_sell, cg and bicgstab samples call, and routes cusparseSpMV /
// cusparseSpMV_bufferSize / cusparseSpMV_preprocess to the hand-written sm_100a kernels (b200spmv_*).
//
// Descriptors stay REAL cuSPARSE descriptors (created by the real library reached through dlopen), so they remain
// valid for everything the samples hand them to afterwards -- cusparseSpSV_* on matL (cg_example.c:392-402,168-181),
// cusparseSpMatSetAttribute (cg_example.c:396-402), cusparseDcsric02 ... -- while a side table remembers what our
// kernels need (pointers, sizes, index base, SELL slice size: CUDA 12.9 has no cusparseSlicedEllGet).
//
// Anything we do not implement (complex / 16-bit / integer value types, CSC/BSR/BlockedELL, CUSPARSE_SPMV_COO_ALG2) is
// forwarded to the real library, so nothing regresses.  B200SPMV_FORWARD=1 forwards everything (A/B runs with one binary).
#include <cuda_runtime_api.h>
#include <cusparse.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/b200spmv.h"
#include "config.h"

namespace {

// ------------------------------------------------------------------------------------------------
// the real library
// ------------------------------------------------------------------------------------------------
struct Real {
    void* h = nullptr;
#define REAL_FN(name) decltype(&::name) name = nullptr;
    REAL_FN(cusparseCreateCsr)
    REAL_FN(cusparseCreateConstCsr)
    REAL_FN(cusparseCreateCoo)
    REAL_FN(cusparseCreateConstCoo)
    REAL_FN(cusparseCreateSlicedEll)
    REAL_FN(cusparseCreateConstSlicedEll)
    REAL_FN(cusparseDestroySpMat)
    REAL_FN(cusparseCsrSetPointers)
    REAL_FN(cusparseCooSetPointers)
    REAL_FN(cusparseSpMatSetValues)
    REAL_FN(cusparseCreateDnVec)
    REAL_FN(cusparseCreateConstDnVec)
    REAL_FN(cusparseDestroyDnVec)
    REAL_FN(cusparseDnVecSetValues)
    REAL_FN(cusparseSpMV_bufferSize)
    REAL_FN(cusparseSpMV_preprocess)
    REAL_FN(cusparseSpMV)
    REAL_FN(cusparseGetStream)
    REAL_FN(cusparseGetPointerMode)
    REAL_FN(cusparseSpMatGetFormat)
    REAL_FN(cusparseConstCsrGet)
    REAL_FN(cusparseConstCooGet)
    REAL_FN(cusparseConstDnVecGet)
    REAL_FN(cusparseSpMM_bufferSize)
    REAL_FN(cusparseSpMM_preprocess)
    REAL_FN(cusparseSpMM)
    REAL_FN(cusparseConstDnMatGet)
    REAL_FN(cusparseDnMatGetStridedBatch)
    REAL_FN(cusparseCsrSetStridedBatch)
    REAL_FN(cusparseSpMatGetStridedBatch)
#undef REAL_FN
    bool forward = false, log = false;
};

Real*          g_real = nullptr;
std::once_flag g_once;

void init_real() {
    Real* r = new Real();
    const char* fwd = getenv("B200SPMV_FORWARD");
    const char* lg = getenv("B200SPMV_LOG");
    r->forward = fwd && fwd[0] && fwd[0] != '0';
    r->log = lg && lg[0] && lg[0] != '0';
    const char* path = getenv("B200SPMV_CUSPARSE");
    const char* cands[] = {path, "libcusparse.so.12", "/usr/local/cuda/lib64/libcusparse.so.12", "libcusparse.so"};
    for (const char* c : cands) {
        if (!c || !c[0]) continue;
        r->h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
        if (r->h) break;
    }
    if (!r->h) {
        fprintf(stderr, "[b200spmv] FATAL: cannot dlopen the real libcusparse.so.12 (%s); set B200SPMV_CUSPARSE\n", dlerror());
        abort();
    }
#define LOAD(name)                                                                                   \
    r->name = (decltype(r->name))dlsym(r->h, #name);                                                 \
    if (!r->name) { fprintf(stderr, "[b200spmv] FATAL: real libcusparse lacks %s\n", #name); abort(); }
    LOAD(cusparseCreateCsr) LOAD(cusparseCreateConstCsr) LOAD(cusparseCreateCoo) LOAD(cusparseCreateConstCoo)
    LOAD(cusparseCreateSlicedEll) LOAD(cusparseCreateConstSlicedEll) LOAD(cusparseDestroySpMat)
    LOAD(cusparseCsrSetPointers) LOAD(cusparseCooSetPointers) LOAD(cusparseSpMatSetValues) LOAD(cusparseCreateDnVec)
    LOAD(cusparseCreateConstDnVec) LOAD(cusparseDestroyDnVec) LOAD(cusparseDnVecSetValues)
    LOAD(cusparseSpMV_bufferSize) LOAD(cusparseSpMV_preprocess) LOAD(cusparseSpMV) LOAD(cusparseGetStream)
    LOAD(cusparseGetPointerMode) LOAD(cusparseSpMatGetFormat) LOAD(cusparseConstCsrGet) LOAD(cusparseConstCooGet)
    LOAD(cusparseConstDnVecGet) LOAD(cusparseSpMM_bufferSize) LOAD(cusparseSpMM_preprocess) LOAD(cusparseSpMM)
    LOAD(cusparseConstDnMatGet) LOAD(cusparseDnMatGetStridedBatch) LOAD(cusparseCsrSetStridedBatch)
    LOAD(cusparseSpMatGetStridedBatch)
#undef LOAD
    g_real = r;
}

inline Real& real() {
    std::call_once(g_once, init_real);
    return *g_real;
}

// ------------------------------------------------------------------------------------------------
// side tables
// ------------------------------------------------------------------------------------------------
struct MatInfo {
    uint64_t             uid = 0;
    cusparseFormat_t     format = CUSPARSE_FORMAT_CSR;
    int64_t              rows = 0, cols = 0, nnz = 0;
    const void *         offsets = nullptr, *row_ind = nullptr, *col_ind = nullptr, *values = nullptr;
    cusparseIndexType_t  off_type = CUSPARSE_INDEX_32I, col_type = CUSPARSE_INDEX_32I;
    cusparseIndexBase_t  base = CUSPARSE_INDEX_BASE_ZERO;
    cudaDataType         vtype = CUDA_R_32F;
    int64_t              sell_values_size = 0, slice_size = 0;
    bool                 use_flat = false;       // preprocess built the flat plan and the row statistic favours csr_flat_kernel
    bool                 use_short = false;      // preprocess found no row longer than b200spmv_csr_short_max_row(): csr_short_kernel
    void*                plan_buffer = nullptr;  // externalBuffer holding this matrix' CSR plan: set ONLY by cusparseSpMV_preprocess
    int                  batch = 1;              // cusparseCsrSetStridedBatch (spmm_csr_batched_example.c:140): matrices in the batch,
    int64_t              off_stride = 0, colval_stride = 0;   //   element strides of the offsets / of the columns and values (0 = shared)
};
struct VecInfo {
    int64_t      size = 0;
    const void*  values = nullptr;
    cudaDataType vtype = CUDA_R_32F;
};

std::mutex                                  g_mu;
std::unordered_map<const void*, MatInfo>    g_mats;
std::unordered_map<const void*, VecInfo>    g_vecs;
std::unordered_map<const void*, uint64_t>   g_plan_owner;  // externalBuffer -> uid of the matrix whose plan it holds
std::atomic<uint64_t>                       g_uid{1};

bool find_mat(cusparseConstSpMatDescr_t d, MatInfo* out) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) { *out = it->second; return true; }
    }
    // Descriptor created behind our back (e.g. by a library that bound the real symbol directly): ask the real getters.
    Real& R = real();
    cusparseFormat_t f;
    if (R.cusparseSpMatGetFormat(d, &f) != CUSPARSE_STATUS_SUCCESS) return false;
    MatInfo m;
    m.format = f;
    if (f == CUSPARSE_FORMAT_CSR) {
        if (R.cusparseConstCsrGet(d, &m.rows, &m.cols, &m.nnz, &m.offsets, &m.col_ind, &m.values, &m.off_type, &m.col_type,
                                  &m.base, &m.vtype) != CUSPARSE_STATUS_SUCCESS)
            return false;
    } else if (f == CUSPARSE_FORMAT_COO) {
        if (R.cusparseConstCooGet(d, &m.rows, &m.cols, &m.nnz, &m.row_ind, &m.col_ind, &m.values, &m.col_type, &m.base,
                                  &m.vtype) != CUSPARSE_STATUS_SUCCESS)
            return false;
        m.off_type = m.col_type;
    } else {
        return false;
    }
    m.uid = g_uid++;
    std::lock_guard<std::mutex> lk(g_mu);
    g_mats[(const void*)d] = m;
    *out = m;
    return true;
}

bool find_vec(cusparseConstDnVecDescr_t d, VecInfo* out) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_vecs.find((const void*)d);
        if (it != g_vecs.end()) { *out = it->second; return true; }
    }
    VecInfo v;
    if (real().cusparseConstDnVecGet(d, &v.size, &v.values, &v.vtype) != CUSPARSE_STATUS_SUCCESS) return false;
    *out = v;
    return true;
}

void record_mat(const void* d, const MatInfo& m) {
    std::lock_guard<std::mutex> lk(g_mu);
    MatInfo mm = m;
    mm.uid = g_uid++;
    g_mats[d] = mm;
}

inline int dtype_of(cudaDataType t) { return t == CUDA_R_32F ? 0 : (t == CUDA_R_64F ? 1 : -1); }

// Which of our kernels take this call?  FAST: the specialised 32-bit-index, single-type kernels.  GENERIC: the plain kernels of
// spmv_generic.cu (64-bit indices, fp32 A with fp64 x / y / arithmetic, Sliced-ELL transposes).  FORWARD: the closed library.
// (COO / Sliced-ELL on the generic kernels: only with B200SPMV_GENERIC=all, see config.h.)
enum Path { FORWARD = 0, FAST = 1, GENERIC = 2 };

inline bool idx_ok(cusparseIndexType_t t) { return t == CUSPARSE_INDEX_32I || t == CUSPARSE_INDEX_64I; }

Path classify(cusparseOperation_t op, const MatInfo& m, const VecInfo& x, const VecInfo& y, cudaDataType compute, cusparseSpMVAlg_t alg) {
    if (m.format != CUSPARSE_FORMAT_CSR && m.format != CUSPARSE_FORMAT_COO && m.format != CUSPARSE_FORMAT_SLICED_ELLPACK) return FORWARD;
    // CUSPARSE_SPMV_COO_ALG2 promises bit-wise reproducible results (cusparse.h:5668-5677, cusparseSpMVAlg_t); our COO kernels add runs that
    // cross warps with floating-point atomics, so that request stays with the closed library.  (CSR / SELL kernels here are
    // reproducible for every alg value.)
    if (m.format == CUSPARSE_FORMAT_COO && alg == CUSPARSE_SPMV_COO_ALG2) return FORWARD;
    const int a_dt = dtype_of(m.vtype), xy_dt = dtype_of(x.vtype);
    if (a_dt < 0 || xy_dt < 0 || y.vtype != x.vtype || compute != x.vtype) return FORWARD;   // complex, 16-bit, integer types
    const bool uniform = a_dt == xy_dt;
    const bool idx32 = m.off_type == CUSPARSE_INDEX_32I && m.col_type == CUSPARSE_INDEX_32I;
    const bool fits32 = m.rows < INT32_MAX && m.cols < INT32_MAX && m.nnz < INT32_MAX - 65536;
    // A^T (== A^H for the real types served here): specialised for CSR (csr_transpose_kernel) and COO (the COO kernel with the
    // index arrays swapped); Sliced-ELL transposes run on the generic kernel
    const bool sell_t = op != CUSPARSE_OPERATION_NON_TRANSPOSE && m.format == CUSPARSE_FORMAT_SLICED_ELLPACK;
    if (uniform && idx32 && fits32 && !sell_t) return FAST;
    const int generic = b200::config().generic;      // 0 off, 1 CSR only (validated on hardware), 2 also COO / Sliced-ELL
    if (generic == 0 || (generic == 1 && m.format != CUSPARSE_FORMAT_CSR)) return FORWARD;
    if (a_dt > xy_dt) return FORWARD;                                   // fp64 A with fp32 vectors: not a cuSPARSE combination
    if (!idx_ok(m.off_type) || !idx_ok(m.col_type)) return FORWARD;
    if (m.off_type == CUSPARSE_INDEX_32I && m.col_type == CUSPARSE_INDEX_64I) return FORWARD;
    return GENERIC;
}

// The call on the kernels of spmv_generic.cu (no plan, no workspace).
int generic_mv(cudaStream_t stream, cusparseOperation_t op, const MatInfo& m, const VecInfo& x, const VecInfo& y, const void* alpha,
               const void* beta, int on_dev) {
    const int tr = op != CUSPARSE_OPERATION_NON_TRANSPOSE;
    const int off64 = m.off_type == CUSPARSE_INDEX_64I, col64 = m.col_type == CUSPARSE_INDEX_64I;
    const int a_dt = dtype_of(m.vtype), xy_dt = dtype_of(x.vtype);
    if (m.format == CUSPARSE_FORMAT_CSR)
        return b200spmv_csr_generic_mv((void*)stream, off64, col64, a_dt, xy_dt, tr, m.rows, m.cols, m.nnz, m.offsets, m.col_ind, m.values,
                                       (int64_t)m.base, alpha, beta, on_dev, x.values, (void*)y.values);
    if (m.format == CUSPARSE_FORMAT_COO)        // A^T: the same entry list with the index arrays (and the shape) swapped
        return tr ? b200spmv_coo_generic_mv((void*)stream, col64, a_dt, xy_dt, m.cols, m.rows, m.nnz, m.col_ind, m.row_ind, m.values,
                                            (int64_t)m.base, alpha, beta, on_dev, x.values, (void*)y.values)
                  : b200spmv_coo_generic_mv((void*)stream, col64, a_dt, xy_dt, m.rows, m.cols, m.nnz, m.row_ind, m.col_ind, m.values,
                                            (int64_t)m.base, alpha, beta, on_dev, x.values, (void*)y.values);
    return b200spmv_sell_generic_mv((void*)stream, off64, col64, a_dt, xy_dt, tr, m.rows, m.cols, m.slice_size, m.offsets, m.col_ind,
                                    m.values, (int64_t)m.base, alpha, beta, on_dev, x.values, (void*)y.values);
}

cusparseStatus_t to_status(int rc) {
    if (rc == 0) return CUSPARSE_STATUS_SUCCESS;
    if (rc == -1) return CUSPARSE_STATUS_INVALID_VALUE;
    return CUSPARSE_STATUS_EXECUTION_FAILED;
}

void logf(const char* what, const MatInfo& m) {
    if (real().log)
        fprintf(stderr, "[b200spmv] %s fmt=%d rows=%lld cols=%lld nnz=%lld vtype=%d base=%d\n", what, (int)m.format,
                (long long)m.rows, (long long)m.cols, (long long)m.nnz, (int)m.vtype, (int)m.base);
}

}  // namespace

// Layout of the caller's externalBuffer for CSR: [tile plan | flat plan (only for matrices that can profit: >= 8 nnz/row)].
static bool flat_eligible(const MatInfo& m) {
    const int mode = b200::config().flat;          // on: every CSR matrix with non-zeros; auto: only where long rows are possible
    return m.nnz > 0 && mode != 0 && (mode == 1 || m.nnz >= 8 * m.rows);
}
static size_t flat_plan_offset(const MatInfo& m) { return (b200spmv_csr_workspace_bytes(m.rows, m.nnz) + 255) / 256 * 256; }
static size_t csr_plans_bytes(const MatInfo& m) {
    return flat_eligible(m) ? flat_plan_offset(m) + b200spmv_csr_flat_workspace_bytes(m.rows, m.nnz)
                            : b200spmv_csr_workspace_bytes(m.rows, m.nnz);
}
// ... then one 256-byte slot for the statistics preprocess reads back (longest row)
static size_t csr_stat_offset(const MatInfo& m) { return (csr_plans_bytes(m) + 255) / 256 * 256; }
static size_t csr_buffer_bytes(const MatInfo& m) { return csr_stat_offset(m) + 256; }

extern "C" {

const char* b200spmv_version(void) { return "b200spmv 0.1 (sm_100a)"; }

// Which path would cusparseSpMV take for a call of this shape?  0: handed to the closed library, 1: the specialised kernels,
// 2: the kernels of spmv_generic.cu.  Pure host logic (no CUDA call, no descriptor): the dispatch table under a CPU test.
int b200spmv_route(int format, int op, int alg, int off_type, int col_type, int a_vtype, int x_vtype, int y_vtype, int compute_type,
                   int64_t rows, int64_t cols, int64_t nnz) {
    MatInfo m;
    m.format = (cusparseFormat_t)format; m.rows = rows; m.cols = cols; m.nnz = nnz;
    m.off_type = (cusparseIndexType_t)off_type; m.col_type = (cusparseIndexType_t)col_type; m.vtype = (cudaDataType)a_vtype;
    VecInfo x, y;
    x.vtype = (cudaDataType)x_vtype; y.vtype = (cudaDataType)y_vtype;
    return (int)classify((cusparseOperation_t)op, m, x, y, (cudaDataType)compute_type, (cusparseSpMVAlg_t)alg);
}

// ---------------------------------------------------------------- sparse-matrix descriptors --------------------------
cusparseStatus_t cusparseCreateCsr(cusparseSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz, void* off, void* col,
                                   void* val, cusparseIndexType_t offT, cusparseIndexType_t colT, cusparseIndexBase_t base,
                                   cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateCsr(d, rows, cols, nnz, off, col, val, offT, colT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_CSR; m.rows = rows; m.cols = cols; m.nnz = nnz; m.offsets = off; m.col_ind = col;
        m.values = val; m.off_type = offT; m.col_type = colT; m.base = base; m.vtype = vT;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseCreateConstCsr(cusparseConstSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz,
                                        const void* off, const void* col, const void* val, cusparseIndexType_t offT,
                                        cusparseIndexType_t colT, cusparseIndexBase_t base, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateConstCsr(d, rows, cols, nnz, off, col, val, offT, colT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_CSR; m.rows = rows; m.cols = cols; m.nnz = nnz; m.offsets = off; m.col_ind = col;
        m.values = val; m.off_type = offT; m.col_type = colT; m.base = base; m.vtype = vT;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseCreateCoo(cusparseSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz, void* row, void* col,
                                   void* val, cusparseIndexType_t idxT, cusparseIndexBase_t base, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateCoo(d, rows, cols, nnz, row, col, val, idxT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_COO; m.rows = rows; m.cols = cols; m.nnz = nnz; m.row_ind = row; m.col_ind = col;
        m.values = val; m.off_type = idxT; m.col_type = idxT; m.base = base; m.vtype = vT;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseCreateConstCoo(cusparseConstSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz,
                                        const void* row, const void* col, const void* val, cusparseIndexType_t idxT,
                                        cusparseIndexBase_t base, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateConstCoo(d, rows, cols, nnz, row, col, val, idxT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_COO; m.rows = rows; m.cols = cols; m.nnz = nnz; m.row_ind = row; m.col_ind = col;
        m.values = val; m.off_type = idxT; m.col_type = idxT; m.base = base; m.vtype = vT;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseCreateSlicedEll(cusparseSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz,
                                         int64_t valuesSize, int64_t sliceSize, void* sliceOff, void* col, void* val,
                                         cusparseIndexType_t offT, cusparseIndexType_t colT, cusparseIndexBase_t base,
                                         cudaDataType vT) {
    cusparseStatus_t st =
        real().cusparseCreateSlicedEll(d, rows, cols, nnz, valuesSize, sliceSize, sliceOff, col, val, offT, colT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_SLICED_ELLPACK; m.rows = rows; m.cols = cols; m.nnz = nnz; m.offsets = sliceOff;
        m.col_ind = col; m.values = val; m.off_type = offT; m.col_type = colT; m.base = base; m.vtype = vT;
        m.sell_values_size = valuesSize; m.slice_size = sliceSize;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseCreateConstSlicedEll(cusparseConstSpMatDescr_t* d, int64_t rows, int64_t cols, int64_t nnz,
                                              int64_t valuesSize, int64_t sliceSize, const void* sliceOff, const void* col,
                                              const void* val, cusparseIndexType_t offT, cusparseIndexType_t colT,
                                              cusparseIndexBase_t base, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateConstSlicedEll(d, rows, cols, nnz, valuesSize, sliceSize, sliceOff, col, val,
                                                              offT, colT, base, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        MatInfo m;
        m.format = CUSPARSE_FORMAT_SLICED_ELLPACK; m.rows = rows; m.cols = cols; m.nnz = nnz; m.offsets = sliceOff;
        m.col_ind = col; m.values = val; m.off_type = offT; m.col_type = colT; m.base = base; m.vtype = vT;
        m.sell_values_size = valuesSize; m.slice_size = sliceSize;
        record_mat((const void*)*d, m);
    }
    return st;
}

cusparseStatus_t cusparseDestroySpMat(cusparseConstSpMatDescr_t d) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) {
            if (it->second.plan_buffer) {
                auto ow = g_plan_owner.find(it->second.plan_buffer);
                if (ow != g_plan_owner.end() && ow->second == it->second.uid) g_plan_owner.erase(ow);
            }
            g_mats.erase(it);
        }
    }
    return real().cusparseDestroySpMat(d);
}

cusparseStatus_t cusparseCsrSetPointers(cusparseSpMatDescr_t d, void* off, void* col, void* val) {
    cusparseStatus_t st = real().cusparseCsrSetPointers(d, off, col, val);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) {
            it->second.offsets = off; it->second.col_ind = col; it->second.values = val;
            it->second.plan_buffer = nullptr;  // structure may have changed: re-analyse on the next SpMV
            it->second.use_flat = false;
            it->second.use_short = false;
        }
    }
    return st;
}

cusparseStatus_t cusparseCooSetPointers(cusparseSpMatDescr_t d, void* row, void* col, void* val) {
    cusparseStatus_t st = real().cusparseCooSetPointers(d, row, col, val);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) { it->second.row_ind = row; it->second.col_ind = col; it->second.values = val; }
    }
    return st;
}

cusparseStatus_t cusparseSpMatSetValues(cusparseSpMatDescr_t d, void* val) {
    cusparseStatus_t st = real().cusparseSpMatSetValues(d, val);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) it->second.values = val;  // the plan is structure-only: still valid
    }
    return st;
}

// cusparse.h:5175 -- spmm_csr_batched_example.c:140.  The real descriptor keeps the setting (forwarded calls see it); the side
// table remembers the strides, which the library offers no getter for.
cusparseStatus_t cusparseCsrSetStridedBatch(cusparseSpMatDescr_t d, int batchCount, int64_t offsetsBatchStride,
                                            int64_t columnsValuesBatchStride) {
    cusparseStatus_t st = real().cusparseCsrSetStridedBatch(d, batchCount, offsetsBatchStride, columnsValuesBatchStride);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)d);
        if (it != g_mats.end()) {
            it->second.batch = batchCount; it->second.off_stride = offsetsBatchStride; it->second.colval_stride = columnsValuesBatchStride;
        }
    }
    return st;
}

// ---------------------------------------------------------------- dense-vector descriptors ---------------------------
cusparseStatus_t cusparseCreateDnVec(cusparseDnVecDescr_t* d, int64_t size, void* values, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateDnVec(d, size, values, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        VecInfo v; v.size = size; v.values = values; v.vtype = vT;
        g_vecs[(const void*)*d] = v;
    }
    return st;
}

cusparseStatus_t cusparseCreateConstDnVec(cusparseConstDnVecDescr_t* d, int64_t size, const void* values, cudaDataType vT) {
    cusparseStatus_t st = real().cusparseCreateConstDnVec(d, size, values, vT);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        VecInfo v; v.size = size; v.values = values; v.vtype = vT;
        g_vecs[(const void*)*d] = v;
    }
    return st;
}

cusparseStatus_t cusparseDestroyDnVec(cusparseConstDnVecDescr_t d) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_vecs.erase((const void*)d);
    }
    return real().cusparseDestroyDnVec(d);
}

cusparseStatus_t cusparseDnVecSetValues(cusparseDnVecDescr_t d, void* values) {
    cusparseStatus_t st = real().cusparseDnVecSetValues(d, values);
    if (st == CUSPARSE_STATUS_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_vecs.find((const void*)d);
        if (it != g_vecs.end()) it->second.values = values;
    }
    return st;
}

// ---------------------------------------------------------------- SpMV -----------------------------------------------
cusparseStatus_t cusparseSpMV_bufferSize(cusparseHandle_t handle, cusparseOperation_t opA, const void* alpha,
                                         cusparseConstSpMatDescr_t matA, cusparseConstDnVecDescr_t vecX, const void* beta,
                                         cusparseDnVecDescr_t vecY, cudaDataType computeType, cusparseSpMVAlg_t alg,
                                         size_t* bufferSize) {
    Real& R = real();
    if (R.forward) return R.cusparseSpMV_bufferSize(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, bufferSize);
    if (!handle || !matA || !vecX || !vecY || !bufferSize) return CUSPARSE_STATUS_INVALID_VALUE;
    // Always ask the real library too: if a later call has to be forwarded, the caller's buffer must be big enough.
    size_t real_size = 0;
    cusparseStatus_t st = R.cusparseSpMV_bufferSize(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, &real_size);
    MatInfo m; VecInfo x, y;
    if (!find_mat(matA, &m) || !find_vec(vecX, &x) || !find_vec(vecY, &y) || classify(opA, m, x, y, computeType, alg) != FAST) {
        *bufferSize = real_size;                   // forwarded calls need the real size; the generic kernels need nothing
        return st;
    }
    if (st != CUSPARSE_STATUS_SUCCESS) return st;  // the real library rejected the arguments: keep its verdict
    size_t ours = 0;
    if (m.format == CUSPARSE_FORMAT_CSR) ours = csr_buffer_bytes(m);
    else if (m.format == CUSPARSE_FORMAT_COO) ours = b200spmv_coo_workspace_bytes(m.rows, m.nnz);
    else ours = b200spmv_sell_workspace_bytes(m.rows, m.sell_values_size, m.slice_size);
    *bufferSize = ours > real_size ? ours : real_size;
    {
        // A fresh bufferSize query usually precedes a fresh cudaMalloc: never trust an older plan after it.
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mats.find((const void*)matA);
        if (it != g_mats.end()) it->second.plan_buffer = nullptr;
    }
    return CUSPARSE_STATUS_SUCCESS;
}

// The CSR tile plan lives in the caller's externalBuffer.  It is TRUSTED on a later cusparseSpMV only if that buffer
// was handed to cusparseSpMV_preprocess for this very descriptor (the documented contract: the buffer must then be
// kept, unmodified, and passed to cusparseSpMV).  Without a preprocess call the buffer is plain scratch to the real
// library -- a caller may share it with SpSV / SpMM or get the same address back from a caching allocator with other
// contents -- so the plan is rebuilt on EVERY call (one ~3 us partition kernel on the same stream, graph-capturable).
static cusparseStatus_t build_csr_plan(cudaStream_t stream, const MatInfo& m, void* buffer) {
    b200::stats().analyze_calls++;
    return to_status(b200spmv_csr_analyze((void*)stream, m.rows, m.nnz, m.offsets, (int32_t)m.base, buffer));
}

static bool plan_is_trusted(const MatInfo& m, void* buffer) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (m.plan_buffer != buffer) return false;
    auto ow = g_plan_owner.find(buffer);
    return ow != g_plan_owner.end() && ow->second == m.uid;
}

cusparseStatus_t cusparseSpMV_preprocess(cusparseHandle_t handle, cusparseOperation_t opA, const void* alpha,
                                         cusparseConstSpMatDescr_t matA, cusparseConstDnVecDescr_t vecX, const void* beta,
                                         cusparseDnVecDescr_t vecY, cudaDataType computeType, cusparseSpMVAlg_t alg,
                                         void* externalBuffer) {
    Real& R = real();
    if (R.forward) return R.cusparseSpMV_preprocess(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer);
    if (!handle || !matA || !vecX || !vecY) return CUSPARSE_STATUS_INVALID_VALUE;
    MatInfo m; VecInfo x, y;
    Path path = FORWARD;
    if (!find_mat(matA, &m) || !find_vec(vecX, &x) || !find_vec(vecY, &y) || (path = classify(opA, m, x, y, computeType, alg)) == FORWARD)
        return R.cusparseSpMV_preprocess(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer);
    const bool tr = opA != CUSPARSE_OPERATION_NON_TRANSPOSE;
    if (x.size != (tr ? m.rows : m.cols) || y.size != (tr ? m.cols : m.rows)) return CUSPARSE_STATUS_INVALID_VALUE;
    if (path == GENERIC || m.format != CUSPARSE_FORMAT_CSR || tr) return CUSPARSE_STATUS_SUCCESS;  // generic kernels / COO / SELL / A^T: no analysis
    if (!externalBuffer || ((uintptr_t)externalBuffer & 15)) return CUSPARSE_STATUS_INVALID_VALUE;
    cudaStream_t stream = nullptr;
    cusparseStatus_t st = R.cusparseGetStream(handle, &stream);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    logf("preprocess(csr plan)", m);
    st = build_csr_plan(stream, m, externalBuffer);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    // The flat plan, and the decision whether this matrix runs on csr_flat_kernel: the share of 32-non-zero steps in
    // which no row ends (R-MAT 1M: 69 %, uniform 16 per row or stencils: 0 %).  Reading the statistic back synchronises
    // the stream once, here in preprocess; while the stream is being captured the read-back is skipped and the tile
    // kernels stay in charge.
    bool use_flat = false, use_short = false;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
    if (flat_eligible(m)) {
        if (cap == cudaStreamCaptureStatusNone) {
            char* fws = (char*)externalBuffer + flat_plan_offset(m);
            int rc = b200spmv_csr_flat_analyze((void*)stream, m.rows, m.nnz, m.offsets, (int32_t)m.base, fws);
            if (rc != 0) return to_status(rc);
            size_t o_ctl = 0;
            b200spmv_csr_flat_plan_offsets(m.rows, m.nnz, nullptr, nullptr, nullptr, &o_ctl);
            int ctl[4] = {0, 0, 0, 0};
            if (cudaMemcpyAsync(ctl, fws + o_ctl, sizeof ctl, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                cudaStreamSynchronize(stream) != cudaSuccess)
                return CUSPARSE_STATUS_EXECUTION_FAILED;
            const int mode = b200::config().flat;
            use_flat = mode == 1 || (mode < 0 && ctl[2] > 0 &&
                                     (long long)ctl[1] * 1000 >= (long long)b200::config().flat_quiet_permille * ctl[2]);
            if (R.log) fprintf(stderr, "[b200spmv] flat plan: %d non-empty rows, %d of %d steps end no row -> %s\n", ctl[0], ctl[1],
                               ctl[2], use_flat ? "csr_flat_kernel" : "tile kernels");
        }
    }
    // All rows short (stencils, meshes)?  The longest row decides; same one-time read-back as above.
    const int short_mode = b200::config().short_rows;
    if (!use_flat && short_mode != 0 && m.nnz > 0 && cap == cudaStreamCaptureStatusNone) {
        int32_t* stat = (int32_t*)((char*)externalBuffer + csr_stat_offset(m));
        int rc = b200spmv_csr_max_row_length((void*)stream, m.rows, m.offsets, stat);
        if (rc != 0) return to_status(rc);
        int32_t longest = 0;
        if (cudaMemcpyAsync(&longest, stat, sizeof longest, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess)
            return CUSPARSE_STATUS_EXECUTION_FAILED;
        use_short = short_mode == 1 || longest <= b200spmv_csr_short_max_row();
        if (R.log) fprintf(stderr, "[b200spmv] longest row: %d non-zeros -> %s\n", (int)longest, use_short ? "csr_short_kernel" : "tile kernels");
    }
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_mats.find((const void*)matA);
    if (it != g_mats.end()) {
        it->second.use_flat = use_flat;
        it->second.use_short = use_short;
        it->second.plan_buffer = externalBuffer;
        g_plan_owner[externalBuffer] = it->second.uid;     // a buffer holds one matrix' plan: the latest preprocess wins
    }
    return CUSPARSE_STATUS_SUCCESS;
}

cusparseStatus_t cusparseSpMV(cusparseHandle_t handle, cusparseOperation_t opA, const void* alpha,
                              cusparseConstSpMatDescr_t matA, cusparseConstDnVecDescr_t vecX, const void* beta,
                              cusparseDnVecDescr_t vecY, cudaDataType computeType, cusparseSpMVAlg_t alg,
                              void* externalBuffer) {
    Real& R = real();
    if (R.forward) {
        b200::stats().forwarded_calls++;
        return R.cusparseSpMV(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer);
    }
    if (!handle || !matA || !vecX || !vecY || !alpha || !beta) return CUSPARSE_STATUS_INVALID_VALUE;
    MatInfo m; VecInfo x, y;
    Path path = FORWARD;
    if (!find_mat(matA, &m) || !find_vec(vecX, &x) || !find_vec(vecY, &y) || (path = classify(opA, m, x, y, computeType, alg)) == FORWARD) {
        if (R.log) fprintf(stderr, "[b200spmv] SpMV forwarded to libcusparse (unsupported combination)\n");
        b200::stats().forwarded_calls++;
        return R.cusparseSpMV(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer);
    }
    const bool tr = opA != CUSPARSE_OPERATION_NON_TRANSPOSE;
    if (x.size != (tr ? m.rows : m.cols) || y.size != (tr ? m.cols : m.rows)) return CUSPARSE_STATUS_INVALID_VALUE;
    cudaStream_t stream = nullptr;
    cusparseStatus_t st = R.cusparseGetStream(handle, &stream);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    cusparsePointerMode_t pm = CUSPARSE_POINTER_MODE_HOST;
    st = R.cusparseGetPointerMode(handle, &pm);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    const int on_dev = pm == CUSPARSE_POINTER_MODE_DEVICE;
    const int dt = dtype_of(m.vtype);
    int rc;
    if (path == GENERIC) {
        logf("SpMV generic kernels (64-bit indices / mixed precision / Sliced-ELL transpose)", m);
        rc = generic_mv(stream, opA, m, x, y, alpha, beta, on_dev);
    } else if (tr && m.format == CUSPARSE_FORMAT_CSR) {
        logf("SpMV csr_transpose_kernel", m);
        rc = b200spmv_csr_transpose_mv((void*)stream, dt, m.rows, m.cols, m.nnz, m.offsets, m.col_ind, m.values, (int32_t)m.base, alpha,
                                       beta, on_dev, x.values, (void*)y.values);
    } else if (tr) {                                        // COO: A^T is the same entry list with the index arrays swapped
        logf("SpMV coo kernel (transposed: index arrays swapped)", m);
        rc = b200spmv_coo_mv((void*)stream, dt, m.cols, m.rows, m.nnz, m.col_ind, m.row_ind, m.values, (int32_t)m.base, alpha,
                             beta, on_dev, x.values, (void*)y.values, externalBuffer);
    } else if (m.format == CUSPARSE_FORMAT_CSR) {
        if (m.rows == 0) return CUSPARSE_STATUS_SUCCESS;
        if (!externalBuffer || ((uintptr_t)externalBuffer & 15)) {
            // No room for a plan (caller ignored bufferSize): the plan-free generic kernel serves it; with that kernel
            // switched off the real library does -- loudly under B200SPMV_LOG, and counted (b200spmv_get_stats) so a test
            // can prove the hot path never takes this exit.
            if (b200::config().generic) {
                logf("SpMV csr_generic_kernel (NULL or misaligned externalBuffer: no room for a plan)", m);
                rc = generic_mv(stream, opA, m, x, y, alpha, beta, on_dev);
                b200::stats().native_calls++;
                return to_status(rc);
            }
            if (R.log) fprintf(stderr, "[b200spmv] SpMV forwarded to libcusparse (NULL or misaligned externalBuffer)\n");
            b200::stats().forwarded_calls++;
            return R.cusparseSpMV(handle, opA, alpha, matA, vecX, beta, vecY, computeType, alg, externalBuffer);
        }
        const bool trusted = plan_is_trusted(m, externalBuffer);
        if (trusted && m.use_flat) {
            logf("SpMV csr_flat_kernel", m);
            rc = b200spmv_csr_flat_mv((void*)stream, dt, m.rows, m.cols, m.nnz, m.offsets, m.col_ind, m.values, (int32_t)m.base, alpha, beta,
                                      on_dev, x.values, (void*)y.values, (char*)externalBuffer + flat_plan_offset(m));
            b200::stats().native_calls++;
            return to_status(rc);
        }
        if (trusted && m.use_short) {
            logf("SpMV csr_short_kernel", m);
            rc = b200spmv_csr_short_mv((void*)stream, dt, m.rows, m.cols, m.nnz, m.offsets, m.col_ind, m.values, (int32_t)m.base, alpha,
                                       beta, on_dev, x.values, (void*)y.values);
            b200::stats().native_calls++;
            return to_status(rc);
        }
        if (!trusted) {
            st = build_csr_plan(stream, m, externalBuffer);
            if (st != CUSPARSE_STATUS_SUCCESS) return st;
        }
        logf("SpMV csr kernels", m);
        rc = b200spmv_csr_mv((void*)stream, dt, m.rows, m.cols, m.nnz, m.offsets, m.col_ind, m.values, (int32_t)m.base, alpha,
                             beta, on_dev, x.values, (void*)y.values, externalBuffer);
    } else if (m.format == CUSPARSE_FORMAT_COO) {
        logf("SpMV coo_tile_kernel", m);
        rc = b200spmv_coo_mv((void*)stream, dt, m.rows, m.cols, m.nnz, m.row_ind, m.col_ind, m.values, (int32_t)m.base, alpha,
                             beta, on_dev, x.values, (void*)y.values, externalBuffer);
    } else {
        logf("SpMV sell_row_kernel", m);
        rc = b200spmv_sell_mv((void*)stream, dt, m.rows, m.cols, m.slice_size, m.offsets, m.col_ind, m.values,
                              (int32_t)m.base, alpha, beta, on_dev, x.values, (void*)y.values, externalBuffer);
    }
    b200::stats().native_calls++;
    return to_status(rc);
}

// ---------------------------------------------------------------- SpMM (CSR x dense) ---------------------------------
// Dense-matrix descriptors are the real library's; what our kernel needs is read through cusparseConstDnMatGet.
struct DnMatInfo {
    int64_t rows = 0, cols = 0, ld = 0;
    const void* values = nullptr;
    cudaDataType vtype = CUDA_R_32F;
    cusparseOrder_t order = CUSPARSE_ORDER_COL;
    int batch = 1;
    int64_t stride = 0;          // elements between consecutive matrices of a strided batch (cusparseDnMatSetStridedBatch)
};
static bool get_dnmat(cusparseConstDnMatDescr_t d, DnMatInfo* m) {
    Real& R = real();
    if (R.cusparseConstDnMatGet(d, &m->rows, &m->cols, &m->ld, &m->values, &m->vtype, &m->order) != CUSPARSE_STATUS_SUCCESS) return false;
    if (R.cusparseDnMatGetStridedBatch(d, &m->batch, &m->stride) != CUSPARSE_STATUS_SUCCESS) { m->batch = 1; m->stride = 0; }
    if (m->batch < 1) m->batch = 1;
    return true;
}
// Strided batches (spmm_csr_batched_example.c:138-160): C_i = alpha * A_i * B_i + beta * C_i for i < N, every operand either
// strided with N entries or shared by the whole batch (A with both strides 0: the sample's "matA broadcast" variant; B with
// batch count 1).  Returns N (1 = no batch), or 0 for a combination we leave to the real library.
static int spmm_batch_count(cusparseConstSpMatDescr_t matA, const MatInfo& a, const DnMatInfo& b, const DnMatInfo& c) {
    const int ab = a.batch < 1 ? 1 : a.batch;
    if (ab == 1 && b.batch == 1 && c.batch == 1) return 1;    // the ordinary call: nothing to ask the library
    if (b200::config().generic < 2) return 0;                 // batches: opt-in until run on hardware (config.h)
    int real_a = 1;
    if (real().cusparseSpMatGetStridedBatch(matA, &real_a) != CUSPARSE_STATUS_SUCCESS || real_a < 1) real_a = 1;
    if (real_a != ab) return 0;                               // the batch was set behind our back: we do not know its strides
    const int n = c.batch;
    if (real_a != 1 && real_a != n) return 0;
    if (b.batch != 1 && b.batch != n) return 0;
    if (n > 1 && c.stride < c.rows * c.cols) return 0;        // overlapping outputs
    return n;
}
static bool spmm_supported(cusparseOperation_t opA, cusparseOperation_t opB, const MatInfo& a, const DnMatInfo& b, const DnMatInfo& c,
                           cudaDataType compute) {
    if (opA != CUSPARSE_OPERATION_NON_TRANSPOSE || opB != CUSPARSE_OPERATION_NON_TRANSPOSE) return false;
    if (a.format != CUSPARSE_FORMAT_CSR || a.off_type != CUSPARSE_INDEX_32I || a.col_type != CUSPARSE_INDEX_32I) return false;
    if (dtype_of(a.vtype) < 0 || b.vtype != a.vtype || c.vtype != a.vtype || compute != a.vtype) return false;
    if (a.rows >= INT32_MAX || a.cols >= INT32_MAX || a.nnz >= INT32_MAX - 65536 || c.cols >= 64 * 65535) return false;
    return true;
}

cusparseStatus_t cusparseSpMM_bufferSize(cusparseHandle_t handle, cusparseOperation_t opA, cusparseOperation_t opB,
                                         const void* alpha, cusparseConstSpMatDescr_t matA, cusparseConstDnMatDescr_t matB,
                                         const void* beta, cusparseDnMatDescr_t matC, cudaDataType computeType,
                                         cusparseSpMMAlg_t alg, size_t* bufferSize) {
    // the real library's answer keeps a forwarded call safe; ours is the row-major copy of a column-major B
    Real& R = real();
    cusparseStatus_t st = R.cusparseSpMM_bufferSize(handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg, bufferSize);
    MatInfo a; DnMatInfo b, c;
    if (st == CUSPARSE_STATUS_SUCCESS && !R.forward && bufferSize && handle && matA && matB && matC && find_mat(matA, &a) &&
        get_dnmat(matB, &b) && get_dnmat(matC, &c) && spmm_supported(opA, opB, a, b, c, computeType) && spmm_batch_count(matA, a, b, c) > 0) {
        const size_t ours = b200spmm_csr_workspace_bytes(dtype_of(a.vtype), b.rows, b.cols, b.order == CUSPARSE_ORDER_ROW);
        if (ours > *bufferSize) *bufferSize = ours;
    }
    return st;
}

cusparseStatus_t cusparseSpMM_preprocess(cusparseHandle_t handle, cusparseOperation_t opA, cusparseOperation_t opB,
                                         const void* alpha, cusparseConstSpMatDescr_t matA, cusparseConstDnMatDescr_t matB,
                                         const void* beta, cusparseDnMatDescr_t matC, cudaDataType computeType,
                                         cusparseSpMMAlg_t alg, void* externalBuffer) {
    Real& R = real();
    MatInfo a; DnMatInfo b, c;
    if (!R.forward && handle && matA && matB && matC && find_mat(matA, &a) && get_dnmat(matB, &b) && get_dnmat(matC, &c) &&
        spmm_supported(opA, opB, a, b, c, computeType) && spmm_batch_count(matA, a, b, c) > 0)
        return CUSPARSE_STATUS_SUCCESS;                              // nothing to analyse
    return R.cusparseSpMM_preprocess(handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg, externalBuffer);
}

cusparseStatus_t cusparseSpMM(cusparseHandle_t handle, cusparseOperation_t opA, cusparseOperation_t opB, const void* alpha,
                              cusparseConstSpMatDescr_t matA, cusparseConstDnMatDescr_t matB, const void* beta,
                              cusparseDnMatDescr_t matC, cudaDataType computeType, cusparseSpMMAlg_t alg, void* externalBuffer) {
    Real& R = real();
    MatInfo a; DnMatInfo b, c;
    int nbatch = 0;
    if (R.forward || !handle || !matA || !matB || !matC || !alpha || !beta || !find_mat(matA, &a) || !get_dnmat(matB, &b) ||
        !get_dnmat(matC, &c) || !spmm_supported(opA, opB, a, b, c, computeType) || (nbatch = spmm_batch_count(matA, a, b, c)) == 0) {
        if (R.log) fprintf(stderr, "[b200spmv] SpMM forwarded to libcusparse\n");
        b200::stats().forwarded_calls++;
        return R.cusparseSpMM(handle, opA, opB, alpha, matA, matB, beta, matC, computeType, alg, externalBuffer);
    }
    if (b.rows != a.cols || c.rows != a.rows || b.cols != c.cols) return CUSPARSE_STATUS_INVALID_VALUE;
    cudaStream_t stream = nullptr;
    cusparseStatus_t st = R.cusparseGetStream(handle, &stream);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    cusparsePointerMode_t pm = CUSPARSE_POINTER_MODE_HOST;
    st = R.cusparseGetPointerMode(handle, &pm);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    if (R.log) fprintf(stderr, "[b200spmv] SpMM spmm_csr_kernel rows=%lld cols=%lld n=%lld nnz=%lld batch=%d\n", (long long)a.rows,
                       (long long)a.cols, (long long)c.cols, (long long)a.nnz, nbatch);
    // externalBuffer: sized by our cusparseSpMM_bufferSize (>= the row-major copy of a column-major B); NULL -> strided walk.
    // A strided batch is one launch sequence per matrix on the handle's stream (the row-major copy of B_i in the buffer is
    // consumed by product i before product i+1 overwrites it: stream order).
    const size_t vsz = a.vtype == CUDA_R_64F ? 8 : 4;
    const int64_t a_off = a.batch > 1 ? a.off_stride : 0, a_cv = a.batch > 1 ? a.colval_stride : 0, b_st = b.batch > 1 ? b.stride : 0;
    int rc = 0;
    for (int i = 0; i < nbatch && rc == 0; i++)
        rc = b200spmm_csr_ws((void*)stream, dtype_of(a.vtype), a.rows, a.cols, c.cols, a.nnz, (const char*)a.offsets + (size_t)i * a_off * 4,
                             (const char*)a.col_ind + (size_t)i * a_cv * 4, (const char*)a.values + (size_t)i * a_cv * vsz, (int32_t)a.base,
                             alpha, beta, pm == CUSPARSE_POINTER_MODE_DEVICE, (const char*)b.values + (size_t)i * b_st * vsz, b.ld,
                             b.order == CUSPARSE_ORDER_ROW, (char*)c.values + (size_t)i * c.stride * vsz, c.ld,
                             c.order == CUSPARSE_ORDER_ROW, externalBuffer);
    b200::stats().native_calls++;
    return to_status(rc);
}

// How many products would cusparseSpMM run on our kernel for these descriptors?  0: the call goes to the closed library (a batch
// layout we do not take, or batches while they are opt-in), 1: an ordinary product, N: a strided batch.  Host logic only (the
// real library's descriptor calls need no device): under a CPU test.
int b200spmm_batch_count(const void* matA, const void* matB, const void* matC) {
    MatInfo a; DnMatInfo b, c;
    if (!matA || !matB || !matC || !find_mat((cusparseConstSpMatDescr_t)matA, &a) || !get_dnmat((cusparseConstDnMatDescr_t)matB, &b) ||
        !get_dnmat((cusparseConstDnMatDescr_t)matC, &c))
        return 0;
    return spmm_batch_count((cusparseConstSpMatDescr_t)matA, a, b, c);
}

}  // extern "C"
