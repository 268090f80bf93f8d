// workload_gen.cu -- device-side synthetic workload generators (bench / test plumbing, not the hot path;
// built into its own libb200gen.so -- the product library libb200spmv.so exports none of this).
//
// Bit-identical to the CPU generators in oracle/spmv_oracle.c (same counter hash, same insertion order), so
// the full-size benchmark matrices never have to cross PCIe and the parity tests can check the integer work
// (edge keys, row offsets, column indices) for exact equality.
//   R-MAT ........ SURVEY.md 8(d) config 2: (a,b,c,d) = (0.57,0.19,0.19,0.05), MSB-first quadrant draws
//   5-pt stencil . cuSPARSE/cg/cg_example.c:71-128, cuSPARSE/bicgstab/bicgstab_example.c:69-127
//   7-pt stencil . cuDSS/simple_residual/laplace_generator.hxx:34-107
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200gen.h"

namespace b200gen {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash2(uint64_t seed_mixed, uint64_t i) { return mix64(seed_mixed + i); }

__global__ void rmat_keys_kernel(uint64_t seed_mixed, int64_t e0, int64_t count, int scale, uint64_t tA, uint64_t tAB,
                                 uint64_t tABC, int64_t rows, int64_t cols, int64_t* __restrict__ keys) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t e = (uint64_t)(e0 + k);
        int64_t r = 0, c = 0;
        for (int l = 0; l < scale; l++) {
            const uint64_t u = hash2(seed_mixed, e * 64u + (uint64_t)l) >> 32;
            const int rb = u >= tAB, cb = (u >= tA && u < tAB) || (u >= tABC);
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        keys[k] = (r < rows && c < cols) ? r * cols + c : (int64_t)-1;
    }
}

template <typename T>
__global__ void uniform_kernel(uint64_t seed_mixed, int64_t i0, int64_t count, T* __restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = hash2(seed_mixed, (uint64_t)(i0 + k));
        out[k] = (T)((double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0);
    }
}

__global__ void stencil5_counts_kernel(int grid, int* __restrict__ counts) {
    const int64_t n = (int64_t)grid * grid;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(row / grid), j = (int)(row % grid);
        counts[row] = 5 - (i == 0) - (i == grid - 1) - (j == 0) - (j == grid - 1);
    }
}

__global__ void stencil5_fill_kernel(int grid, double mass, double ux, double uy, const int* __restrict__ off,
                                     int* __restrict__ col, double* __restrict__ val) {
    const int64_t n = (int64_t)grid * grid;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(row / grid), j = (int)(row % grid);
        int it = off[row];
        if (i > 0)        { col[it] = (int)(row - grid); val[it++] = -1.0 - ux; }
        if (j > 0)        { col[it] = (int)(row - 1);    val[it++] = -1.0 - uy; }
        col[it] = (int)row; val[it++] = 4.0 + mass + ux + uy;
        if (j < grid - 1) { col[it] = (int)(row + 1);    val[it++] = -1.0; }
        if (i < grid - 1) { col[it] = (int)(row + grid); val[it++] = -1.0; }
    }
}

__global__ void laplace7_counts_kernel(int nx, int* __restrict__ counts) {
    const int64_t n = (int64_t)nx * nx * nx;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(row % nx), y = (int)((row / nx) % nx), z = (int)(row / ((int64_t)nx * nx));
        counts[row] = 7 - (z == 0) - (z == nx - 1) - (y == 0) - (y == nx - 1) - (x == 0) - (x == nx - 1);
    }
}

template <typename T>
__global__ void laplace7_fill_kernel(int nx, const int* __restrict__ off, int* __restrict__ col, T* __restrict__ val) {
    const int64_t n = (int64_t)nx * nx * nx, nxy = (int64_t)nx * nx;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(row % nx), y = (int)((row / nx) % nx), z = (int)(row / nxy);
        int it = off[row];
        if (z > 0)      { col[it] = (int)(row - nxy); val[it++] = T(-1); }
        if (y > 0)      { col[it] = (int)(row - nx);  val[it++] = T(-1); }
        if (x > 0)      { col[it] = (int)(row - 1);   val[it++] = T(-1); }
        col[it] = (int)row; val[it++] = T(16);
        if (x < nx - 1) { col[it] = (int)(row + 1);   val[it++] = T(-1); }
        if (y < nx - 1) { col[it] = (int)(row + nx);  val[it++] = T(-1); }
        if (z < nx - 1) { col[it] = (int)(row + nxy); val[it++] = T(-1); }
    }
}

static inline unsigned grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    if (b > 148 * 32) b = 148 * 32;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace b200gen

using namespace b200gen;

extern "C" {

int b200gen_rmat_keys(void* stream, uint64_t seed, int64_t e0, int64_t count, int32_t scale, uint64_t tA, uint64_t tAB,
                      uint64_t tABC, int64_t rows, int64_t cols, int64_t* keys_out) {
    if (count <= 0) return 0;
    rmat_keys_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(mix64(seed), e0, count, scale, tA, tAB, tABC,
                                                                             rows, cols, keys_out);
    return (int)cudaGetLastError();
}

int b200gen_uniform(void* stream, int dtype, uint64_t seed, int64_t i0, int64_t count, void* out) {
    if (count <= 0) return 0;
    if (dtype == 0)
        uniform_kernel<float><<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(mix64(seed), i0, count, (float*)out);
    else if (dtype == 1)
        uniform_kernel<double><<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(mix64(seed), i0, count, (double*)out);
    else
        return -1;
    return (int)cudaGetLastError();
}

int b200gen_stencil5_counts(void* stream, int32_t grid, int32_t* counts_out) {
    stencil5_counts_kernel<<<grid_for((int64_t)grid * grid, 256), 256, 0, (cudaStream_t)stream>>>(grid, counts_out);
    return (int)cudaGetLastError();
}

int b200gen_stencil5_fill(void* stream, int32_t grid, double mass, double ux, double uy, const int32_t* row_offsets,
                          int32_t* col_out, double* val_out) {
    stencil5_fill_kernel<<<grid_for((int64_t)grid * grid, 256), 256, 0, (cudaStream_t)stream>>>(grid, mass, ux, uy,
                                                                                               row_offsets, col_out, val_out);
    return (int)cudaGetLastError();
}

int b200gen_laplace7_counts(void* stream, int32_t nx, int32_t* counts_out) {
    laplace7_counts_kernel<<<grid_for((int64_t)nx * nx * nx, 256), 256, 0, (cudaStream_t)stream>>>(nx, counts_out);
    return (int)cudaGetLastError();
}

int b200gen_laplace7_fill(void* stream, int dtype, int32_t nx, const int32_t* row_offsets, int32_t* col_out, void* val_out) {
    const unsigned g = grid_for((int64_t)nx * nx * nx, 256);
    if (dtype == 0)
        laplace7_fill_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(nx, row_offsets, col_out, (float*)val_out);
    else if (dtype == 1)
        laplace7_fill_kernel<double><<<g, 256, 0, (cudaStream_t)stream>>>(nx, row_offsets, col_out, (double*)val_out);
    else
        return -1;
    return (int)cudaGetLastError();
}

}  // extern "C"
