// cuda_emulation.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to compile a kernel's own source for the host
// and run it without a GPU (tests/test_generic_emulation.py).  Not part of the product; never linked into libb200spmv.so.
//
//   * one std::thread per lane, a warp (32 lanes) at a time, warps and CTAs one after the other: intra-warp lock-step is NOT
//     assumed anywhere -- lanes meet only inside __shfl_down_sync, which is a real barrier between the lanes named in the mask
//     (all 32 here), exactly the guarantee the *_sync intrinsics give on the device;
//   * threadIdx / blockIdx are thread-local, blockDim / gridDim process-wide;
//   * atomicAdd is a locked read-modify-write (lanes of a warp do run concurrently);
//   * Scalars / axpby restate cudalibrarysamples_b200/csrc/spmv_common.cuh:60-75 (that header's streaming loads are inline
//     PTX and cannot be compiled for the host; the kernels emulated here do not use them).
// Limits: no __shared__, no __syncthreads(), one warp in flight -- enough for the plan-free kernels of spmv_generic_kernels.cuh.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

namespace emu {

// a reusable barrier for the 32 lanes of the warp in flight, plus the exchange slots of the shuffle
struct Warp {
    std::mutex              m;
    std::condition_variable cv;
    int                     waiting = 0, generation = 0, lanes = 32;
    alignas(16) unsigned char slot[32][16];
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(m);
        const int gen = generation;
        if (++waiting == lanes) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != generation; });
    }
};
inline Warp*        g_warp = nullptr;
inline std::mutex   g_atomic_mutex;

// Runs kernel() once per thread of a <<<grid, block>>> launch (1-D), a warp of real threads at a time.
inline void launch(unsigned grid, unsigned block, const std::function<void()>& kernel) {
    gridDim.x = grid;
    blockDim.x = block;
    Warp w;
    g_warp = &w;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned w0 = 0; w0 < block; w0 += 32) {
            const unsigned n = block - w0 < 32 ? block - w0 : 32;
            w.lanes = (int)n;
            w.waiting = 0;
            std::vector<std::thread> lanes;
            for (unsigned l = 0; l < n; l++)
                lanes.emplace_back([=, &kernel] {
                    blockIdx.x = b;
                    threadIdx.x = w0 + l;
                    kernel();
                });
            for (auto& t : lanes) t.join();
        }
    g_warp = nullptr;
}

}  // namespace emu

// __shfl_down_sync(mask, v, delta, width): lane i of a width-lane segment reads lane i + delta's value, or keeps its own when
// that lane lies outside the segment (CUDA C++ Programming Guide, warp shuffle functions).  Every lane of the warp calls it.
template <typename T>
inline T __shfl_down_sync(unsigned /*mask: all lanes*/, T v, int delta, int width = 32) {
    static_assert(sizeof(T) <= 16, "shuffle payload");
    emu::Warp& w = *emu::g_warp;
    const int lane = (int)(threadIdx.x & 31);
    std::memcpy(w.slot[lane], &v, sizeof(T));
    w.arrive_and_wait();                                    // every lane has published its value
    const int src = (lane % width) + delta < width ? lane + delta : lane;
    T r;
    std::memcpy(&r, w.slot[src < w.lanes ? src : lane], sizeof(T));
    w.arrive_and_wait();                                    // every lane has read: the slots may be overwritten
    return r;
}

template <typename T>
inline T atomicAdd(T* p, T v) {
    std::lock_guard<std::mutex> lk(emu::g_atomic_mutex);
    const T old = *p;
    *p = old + v;
    return old;
}

namespace b200 {

template <typename T>
struct Scalars {
    T        alpha, beta;
    const T* alpha_dev;
    const T* beta_dev;
    T a() const { return alpha_dev ? *alpha_dev : alpha; }
    T b() const { return beta_dev ? *beta_dev : beta; }
};

template <typename T>
inline T axpby(T alpha, T sum, T beta, const T* y) {
    return beta == T(0) ? alpha * sum : alpha * sum + beta * (*y);
}

}  // namespace b200
