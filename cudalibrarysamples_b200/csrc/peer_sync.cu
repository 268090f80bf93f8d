// peer_sync.cu -- a device-side barrier over NVLink peer memory for the GPUs of one box (one process per GPU).
//
// Used by the sharded SpMV step (cudalibrarysamples_b200/sharded.py): "every rank has written its x shard" must be known on
// every GPU before the copy engines start pulling shards.  torch's symmetric-memory barrier cannot be replayed inside a CUDA
// graph (its sequence state lives on the host: the replayed step of round 2 hung), and a 1-element NCCL all-reduce costs
// ~15 us per step.  This one keeps its epoch in DEVICE memory, so a captured launch stays correct on every replay:
//
//   flags[r]   (one 8-byte slot per rank, in every rank's symmetric buffer)  = the last epoch rank r has announced HERE
//   epoch      (local device memory)                                          = the number of barriers this rank has entered
//
//   barrier:  e = ++epoch;  release-store e into flags[my_rank] of EVERY rank (peer stores over NVLink);
//             spin (acquire loads, local memory) until flags[r] >= e for every r.
//
// One warp; lane r talks to rank r.  A peer that never arrives trips the timeout and the kernel traps (a CUDA error, not a
// hung GPU).  System-scope fences order the data written before the barrier (the x shard) against the flag.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200spmv.h"

namespace b200peer {

__global__ void barrier_kernel(unsigned long long* const* __restrict__ peer_flags, unsigned long long* __restrict__ epoch,
                               int my_rank, int world, long long timeout_cycles) {
    const int lane = (int)threadIdx.x;
    unsigned long long e = 0;
    if (lane == 0) e = *epoch + 1ull;
    e = __shfl_sync(0xffffffffu, e, 0);
    __threadfence_system();                                   // everything this stream wrote before is visible box-wide
    if (lane < world) {
        unsigned long long* dst = peer_flags[lane] + my_rank;  // slot my_rank in rank `lane`'s flag array
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(e) : "memory");
    }
    if (lane < world) {
        const unsigned long long* src = peer_flags[my_rank] + lane;   // my own array: who has announced epoch e here?
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if (v >= e) break;
            if (clock64() - t0 > timeout_cycles) __trap();    // a rank is missing: fail loudly instead of hanging the GPU
        }
    }
    __syncwarp();
    __threadfence_system();
    if (lane == 0) *epoch = e;
}

// dst[0, bytes) = src[0, bytes): 16-byte loads from (peer) memory over NVLink, 16-byte stores to local memory.  An SM copy
// instead of a copy-engine copy: measured at N = 2, one 8 MB cudaMemcpyAsync out of a peer's symmetric buffer took ~40 us
// (~200 GB/s); LDG.128 from peer memory sustains several hundred GB/s with a few dozen CTAs.
__global__ void __launch_bounds__(256) pull_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long n16,
                                                   char* __restrict__ dst_tail, const char* __restrict__ src_tail, int tail) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {              // four independent 16-byte loads in flight per thread
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

}  // namespace b200peer

extern "C" int b200peer_pull(void* stream, void* dst, const void* src, size_t bytes, int ctas) {
    if (!dst || !src || ctas < 1 || (((uintptr_t)dst | (uintptr_t)src) & 15)) return -1;
    if (bytes == 0) return 0;
    const long long n16 = (long long)(bytes / 16);
    const int tail = (int)(bytes % 16);
    b200peer::pull_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>((uint4*)dst, (const uint4*)src, n16, (char*)dst + n16 * 16,
                                                                  (const char*)src + n16 * 16, tail);
    return (int)cudaGetLastError();
}

extern "C" int b200peer_barrier(void* stream, const void* peer_flag_ptrs_dev, void* epoch_dev, int my_rank, int world,
                                double timeout_seconds) {
    if (!peer_flag_ptrs_dev || !epoch_dev || world < 1 || world > 32 || my_rank < 0 || my_rank >= world) return -1;
    const long long cycles = (long long)(timeout_seconds * 1.9e9);
    b200peer::barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)peer_flag_ptrs_dev,
                                                                 (unsigned long long*)epoch_dev, my_rank, world, cycles);
    return (int)cudaGetLastError();
}
