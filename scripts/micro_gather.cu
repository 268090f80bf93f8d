// micro_gather.cu -- how fast can one B200 SM gather 8-byte elements of x at scattered indices?
// Paths: (a) ld.global.nc.f64 through the LSU pipe, (b) tex1Dfetch<int2> through the TEX pipe, (c) LSU with
// L1::no_allocate, (d) 16 active lanes per instruction.  Index streams: uniform random and R-MAT-skewed.
// Round 2 additions (VERDICT r1 "x tiles in shared memory"): (e) the same gathers with the L1 shrunk by a dynamic
// shared-memory allocation (how much does the x gather owe to L1 hits?), (f) gathers out of a shared-memory table
// (LDS.64 at random indices: the bank-conflict-limited rate), (g) a hot/cold split: indices below HOT are served from
// a shared-memory copy of x[0, HOT), the rest through LDG -- the cost model of a "hot columns of x staged in shared
// memory" SpMV.  An index stream can be read from a file (argv[1], int32, e.g. the col_ind[] of the R-MAT matrix in
// CSR order) so the numbers apply to the real access order, not only to synthetic streams.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o micro_gather micro_gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}

__global__ void gen_idx(int* idx, int64_t n, int ncols, int mode, int scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mode == 0) idx[i] = (int)(mix64(i) % (uint64_t)ncols);
        else {
            int c;
            uint64_t k = 0;
            do {
                c = 0;
                for (int l = 0; l < scale; l++) { uint64_t u = mix64(i * 64 + l + (k << 40)) >> 32; c = (c << 1) | (u < (uint64_t)(0.24 * 4294967296.0) ? 1 : 0); }
                k++;
            } while (c >= ncols);
            idx[i] = c;
        }
    }
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const int* __restrict__ idx, const double* __restrict__ x, cudaTextureObject_t tex,
                                                     double* __restrict__ out, int64_t n) {
    double acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        int c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            int v;
            asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(idx + i + u * stride));
            c[u] = v;
        }
        double xv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE == 0) xv[u] = __ldg(x + c[u]);
            else if (MODE == 1) { int2 t = tex1Dfetch<int2>(tex, c[u]); xv[u] = __hiloint2double(t.y, t.x); }
            else if (MODE == 2) { double r; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(x + c[u])); xv[u] = r; }
            else if (MODE == 3) { xv[u] = (threadIdx.x & 1) ? 0.0 : __ldg(x + c[u]); }
            else if (MODE == 4) { double r; asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(x + c[u])); xv[u] = r; }
            else xv[u] = __ldg(x + c[u]);
        }
        if (MODE == 5) {   // + 10 warp shuffles per gathered element (segmented-scan cost model)
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, xv[u], o); if ((threadIdx.x & 31) >= o) xv[u] += t; }
            }
        }
        if (MODE == 6) {   // + one conflict-free STS.64 and LDS.64 per gathered element (products parked in smem)
            __shared__ double sm[256 * UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) sm[u * 256 + threadIdx.x] = xv[u];
            __syncthreads();
#pragma unroll
            for (int u = 0; u < UNROLL; u++) xv[u] = sm[u * 256 + (threadIdx.x ^ 32)];
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += xv[u];
    }
    if (acc == 123.456) out[0] = acc;
}

// MODE 0: every index through LDG (reference, same launch shape).  MODE 1: indices < hot from the shared-memory copy of
// x[0, hot), the others through LDG.  MODE 2: (idx % hot) from shared memory only (pure LDS gather rate).
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(1024) hot_gather_kernel(const int* __restrict__ idx, const double* __restrict__ x,
                                                          double* __restrict__ out, int64_t n, int hot) {
    extern __shared__ double sx[];
    if (MODE != 0) for (int i = threadIdx.x; i < hot; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    double acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        int c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            int v;
            asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(idx + i + u * stride));
            c[u] = v;
        }
        double xv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE == 0) xv[u] = __ldg(x + c[u]);
            else if (MODE == 1) xv[u] = c[u] < hot ? sx[c[u]] : __ldg(x + c[u]);
            else xv[u] = sx[(unsigned)c[u] % (unsigned)hot];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += xv[u];
    }
    if (acc == 123.456) out[0] = acc;
}

template <int MODE>
float run_hot(const char* name, const int* idx, const double* x, double* out, int64_t n, int threads, int hot, size_t dyn) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(hot_gather_kernel<MODE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int per = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, hot_gather_kernel<MODE, 8>, threads, dyn));
    const int blocks = 148 * (per < 1 ? 1 : per);
    for (int w = 0; w < 3; w++) hot_gather_kernel<MODE, 8><<<blocks, threads, dyn>>>(idx, x, out, n, hot);
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int r = 0; r < reps; r++) hot_gather_kernel<MODE, 8><<<blocks, threads, dyn>>>(idx, x, out, n, hot);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("  %-58s %8.2f us  %7.2f Gelem/s  (%.3f elem/clk/SM)  [%d CTAs/SM x %d thr, %zu KB smem/CTA]\n", name, ms * 1e3,
           n / (ms * 1e-3) / 1e9, n / (ms * 1e-3) / 148 / 1.965e9, per, threads, dyn >> 10);
    return ms;
}

// ascending-popcount remap: the R-MAT column distribution is a product of per-bit Bernoulli(0.24) draws, so the hottest
// columns are the ones with the fewest 1 bits.  rank[] maps a column to its position in (popcount, value) order -- what a
// preprocess pass that sorts columns by access count would produce; the remapped stream puts the hot columns at [0, HOT).
__global__ void remap_idx(const int* __restrict__ in, int* __restrict__ outi, const int* __restrict__ rank, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) outi[i] = rank[in[i]];
}

template <int MODE>
float run(const char* name, const int* idx, const double* x, cudaTextureObject_t tex, double* out, int64_t n, int blocks) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) gather_kernel<MODE, 8><<<blocks, 256>>>(idx, x, tex, out, n);
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int r = 0; r < reps; r++) gather_kernel<MODE, 8><<<blocks, 256>>>(idx, x, tex, out, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("  %-34s %8.2f us  %7.2f Gelem/s  (%.3f elem/clk/SM @1.965GHz)\n", name, ms * 1e3, n / (ms * 1e-3) / 1e9,
           n / (ms * 1e-3) / 148 / 1.965e9);
    return ms;
}

int main(int argc, char** argv) {
    int64_t n = 16 * 1000 * 1000;
    int* idx; CK(cudaMalloc(&idx, n * 4));
    double* out; CK(cudaMalloc(&out, 8));
    if (argc > 1) {
        // ---- round 2: real index stream (col_ind[] of the benchmark matrix in CSR order), x = 1M doubles ----
        const int ncols = argc > 2 ? atoi(argv[2]) : 1000000;
        FILE* f = fopen(argv[1], "rb");
        if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
        std::vector<int> h((size_t)n);
        n = (int64_t)fread(h.data(), 4, (size_t)n, f); fclose(f);
        n = n / (256 * 8 * 148) * (256 * 8 * 148);
        CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
        double* x; CK(cudaMalloc(&x, (size_t)ncols * 8)); CK(cudaMemset(x, 0, (size_t)ncols * 8));
        // rank by (popcount, value)
        std::vector<int> order(ncols), rank(ncols);
        for (int i = 0; i < ncols; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [](int a, int b) { return __builtin_popcount(a) < __builtin_popcount(b); });
        for (int i = 0; i < ncols; i++) rank[order[i]] = i;
        int* drank; CK(cudaMalloc(&drank, (size_t)ncols * 4)); CK(cudaMemcpy(drank, rank.data(), (size_t)ncols * 4, cudaMemcpyHostToDevice));
        int* idx2; CK(cudaMalloc(&idx2, n * 4));
        remap_idx<<<148 * 8, 256>>>(idx, idx2, drank, n); CK(cudaDeviceSynchronize());
        printf("index stream from %s: %lld gathers, x = %d doubles\n", argv[1], (long long)n, ncols);
        printf(" (1) how much of the gather rate is L1 hits?  same LDG gathers, L1 shrunk by a dynamic smem allocation\n");
        for (size_t kb : {0, 8, 16, 24, 27})
            run_hot<0>("LDG gathers, CSR order", idx, x, out, n, 256, 0, kb << 10);
        printf(" (2) pure shared-memory gather rate (LDS.64 at idx %% table)\n");
        for (int hot : {4096, 12288, 24576})
            run_hot<2>("LDS gathers only", idx, x, out, n, 1024, hot, (size_t)hot * 8);
        printf(" (3) hot/cold split on the popcount-ranked stream: columns [0, HOT) from shared memory, the rest LDG\n");
        run_hot<0>("ranked stream, all LDG (1024 thr, no smem)", idx2, x, out, n, 1024, 0, 0);
        for (int hot : {4096, 8192, 12288, 16384, 24576}) {
            long long hits = 0;
            for (int64_t i = 0; i < n; i += 97) hits += rank[h[(size_t)i]] < hot;
            char nm[96]; snprintf(nm, sizeof nm, "HOT=%d (%.0f%% of gathers hit smem)", hot, 100.0 * hits / ((n + 96) / 97));
            run_hot<1>(nm, idx2, x, out, n, 1024, hot, (size_t)hot * 8);
        }
        for (int hot : {4096, 8192, 11264}) {
            char nm[96]; snprintf(nm, sizeof nm, "HOT=%d, 512 thr x 2 CTAs/SM", hot);
            run_hot<1>(nm, idx2, x, out, n, 512, hot, (size_t)hot * 8);
        }
        return 0;
    }
    for (int ncols : {1000000}) {
        double* x; CK(cudaMalloc(&x, (size_t)ncols * 8)); CK(cudaMemset(x, 0, (size_t)ncols * 8));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (size_t)ncols * 8;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        int scale = 0; while ((1 << scale) < ncols) scale++;
        for (int mode : {0, 1}) {
            gen_idx<<<148 * 8, 256>>>(idx, n, ncols, mode, scale); CK(cudaDeviceSynchronize());
            printf("x: %d doubles (%.0f MB), indices: %s, 16M gathers, grid 148*8 x 256, unroll 8\n", ncols, ncols * 8e-6,
                   mode ? "R-MAT-skewed (bit=1 w.p. 0.24)" : "uniform random");
            for (int blocks : {148 * 4, 148 * 8}) {
                printf(" blocks=%d\n", blocks);
                run<0>("LSU  ld.global.nc.f64", idx, x, tex, out, n, blocks);
                run<1>("TEX  tex1Dfetch<int2>", idx, x, tex, out, n, blocks);
                run<2>("LSU  ld.global.nc.L1::no_allocate", idx, x, tex, out, n, blocks);
                run<3>("LSU  16 of 32 lanes active (n/2 elems)", idx, x, tex, out, n, blocks);
                run<4>("LSU  ld.global.f64 (coherent)", idx, x, tex, out, n, blocks);
                run<5>("LSU  + 10 SHFL per element", idx, x, tex, out, n, blocks);
                run<6>("LSU  + STS.64 + LDS.64 per element", idx, x, tex, out, n, blocks);
            }
        }
        CK(cudaDestroyTextureObject(tex)); CK(cudaFree(x));
    }
    return 0;
}
