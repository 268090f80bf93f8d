/* spmv_timer.c -- C-side timing harness for the cusparseSpMV path, written in the style of the reference samples
 * (cuSPARSE/spmv_csr/spmv_csr_example.c:86-118: CreateCsr / CreateDnVec x2 / SpMV_bufferSize / cudaMalloc /
 * SpMV_preprocess / SpMV), so that the closed library can be timed as the CUDA-12.9 TOOLKIT build the samples link
 * (libcusparse 12.5.10) -- inside a Python process that imported torch only torch's bundled 12.5.8 can be loaded.
 *
 * Built twice by cudalibrarysamples_b200/build.py:
 *   tools/_bin/spmv_timer.cusparse   gcc ... -lcusparse                 (the closed library = the bar to beat)
 *   tools/_bin/spmv_timer.b200       gcc ... -lb200spmv -lcusparse      (the same binary logic through the shim)
 *
 * usage: spmv_timer <dir> <rows> <cols> <nnz> <steps> <warmup>
 *   <dir>/off.bin (int32 rows+1), col.bin (int32 nnz), val.bin (f64 nnz), x.bin (f64 cols); writes <dir>/y_<tag>.bin
 * prints one JSON line: {"us_per_spmv": median-free mean over <steps> back-to-back calls between two CUDA events, ...}
 */
#include <cuda_runtime_api.h>
#include <cusparse.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
#define CS(x) do { cusparseStatus_t s_ = (x); if (s_ != CUSPARSE_STATUS_SUCCESS) { fprintf(stderr, "cuSPARSE error %d at line %d\n", (int)s_, __LINE__); return 3; } } while (0)

static void* slurp(const char* dir, const char* name, size_t bytes) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(4); }
    void* p = malloc(bytes ? bytes : 1);
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read on %s\n", path); exit(4); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s dir rows cols nnz steps warmup [tag]\n", argv[0]); return 1; }
    const char* dir = argv[1];
    const long long rows = atoll(argv[2]), cols = atoll(argv[3]), nnz = atoll(argv[4]);
    const int steps = atoi(argv[5]), warmup = atoi(argv[6]);
    const char* tag = argc > 7 ? argv[7] : "out";
    int*    h_off = (int*)slurp(dir, "off.bin", (size_t)(rows + 1) * 4);
    int*    h_col = (int*)slurp(dir, "col.bin", (size_t)nnz * 4);
    double* h_val = (double*)slurp(dir, "val.bin", (size_t)nnz * 8);
    double* h_x   = (double*)slurp(dir, "x.bin", (size_t)cols * 8);
    int *d_off, *d_col; double *d_val, *d_x, *d_y;
    CK(cudaMalloc((void**)&d_off, (size_t)(rows + 1) * 4)); CK(cudaMalloc((void**)&d_col, (size_t)nnz * 4));
    CK(cudaMalloc((void**)&d_val, (size_t)nnz * 8)); CK(cudaMalloc((void**)&d_x, (size_t)cols * 8)); CK(cudaMalloc((void**)&d_y, (size_t)rows * 8));
    CK(cudaMemcpy(d_off, h_off, (size_t)(rows + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_col, h_col, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_val, h_val, (size_t)nnz * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, h_x, (size_t)cols * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_y, 0, (size_t)rows * 8));

    cusparseHandle_t handle; cusparseSpMatDescr_t matA; cusparseDnVecDescr_t vecX, vecY;
    const double alpha = 1.0, beta = 0.0;
    size_t bufferSize = 0; void* dBuffer = NULL;
    CS(cusparseCreate(&handle));
    CS(cusparseCreateCsr(&matA, rows, cols, nnz, d_off, d_col, d_val, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                         CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    CS(cusparseCreateDnVec(&vecX, cols, d_x, CUDA_R_64F));
    CS(cusparseCreateDnVec(&vecY, rows, d_y, CUDA_R_64F));
    CS(cusparseSpMV_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_64F,
                               CUSPARSE_SPMV_ALG_DEFAULT, &bufferSize));
    CK(cudaMalloc(&dBuffer, bufferSize ? bufferSize : 16));
    CS(cusparseSpMV_preprocess(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_64F,
                               CUSPARSE_SPMV_ALG_DEFAULT, dBuffer));
    for (int i = 0; i < warmup; i++)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_64F,
                        CUSPARSE_SPMV_ALG_DEFAULT, dBuffer));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < steps; i++)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_64F,
                        CUSPARSE_SPMV_ALG_DEFAULT, dBuffer));
    CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    double* h_y = (double*)malloc((size_t)rows * 8);
    CK(cudaMemcpy(h_y, d_y, (size_t)rows * 8, cudaMemcpyDeviceToHost));
    char path[1024]; snprintf(path, sizeof path, "%s/y_%s.bin", dir, tag);
    FILE* f = fopen(path, "wb"); if (f) { fwrite(h_y, 8, (size_t)rows, f); fclose(f); }
    int ver = 0; cusparseGetVersion(handle, &ver);
    printf("{\"us_per_spmv\": %.3f, \"steps\": %d, \"warmup\": %d, \"cusparse_version\": %d, \"buffer_bytes\": %zu, \"rows\": %lld, \"nnz\": %lld}\n",
           1e3 * ms / steps, steps, warmup, ver, bufferSize, rows, nnz);
    CS(cusparseDestroySpMat(matA)); CS(cusparseDestroyDnVec(vecX)); CS(cusparseDestroyDnVec(vecY)); CS(cusparseDestroy(handle));
    cudaFree(dBuffer); cudaFree(d_off); cudaFree(d_col); cudaFree(d_val); cudaFree(d_x); cudaFree(d_y);
    return 0;
}
