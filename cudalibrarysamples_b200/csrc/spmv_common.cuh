// spmv_common.cuh -- device helpers shared by the CSR / COO / SELL kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ---- streaming loads: val[] / col_ind[] are read exactly once per SpMV, so they bypass L1
//      (ld.global.nc.L1::no_allocate) and leave the whole unified L1 to the gathered x vector. ----
__device__ __forceinline__ int4 ldg_stream_int4(const int* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ldg_stream_int2(const int* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream_v2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_v4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// 4 consecutive values starting at a 16-byte-aligned element index.
__device__ __forceinline__ void load4_stream(const double* p, double (&v)[4]) {
    double2 a = ldg_stream_v2(p), b = ldg_stream_v2(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void load4_stream(const float* p, float (&v)[4]) {
    float4 a = ldg_stream_v4(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// ---- alpha / beta: by value (CUSPARSE_POINTER_MODE_HOST) or read in-kernel from device memory
//      (CUSPARSE_POINTER_MODE_DEVICE, cusparse.h:275-278). ----
template <typename T>
struct Scalars {
    T        alpha, beta;
    const T* alpha_dev;
    const T* beta_dev;
    __device__ __forceinline__ T a() const { return alpha_dev ? *alpha_dev : alpha; }
    __device__ __forceinline__ T b() const { return beta_dev ? *beta_dev : beta; }
};

template <typename T>
__device__ __forceinline__ T axpby(T alpha, T sum, T beta, const T* y) {
    // beta == 0 must not read y (it may hold NaN/uninitialised memory, as cudaMalloc'ed dY does).
    return beta == T(0) ? alpha * sum : alpha * sum + beta * (*y);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace b200
