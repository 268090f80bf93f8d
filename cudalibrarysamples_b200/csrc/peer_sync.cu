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

}  // namespace b200peer

extern "C" int b200peer_barrier(void* stream, const void* peer_flag_ptrs_dev, void* epoch_dev, int my_rank, int world,
                                double timeout_seconds) {
    if (!peer_flag_ptrs_dev || !epoch_dev || world < 1 || world > 32 || my_rank < 0 || my_rank >= world) return -1;
    const long long cycles = (long long)(timeout_seconds * 1.9e9);
    b200peer::barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)peer_flag_ptrs_dev,
                                                                 (unsigned long long*)epoch_dev, my_rank, world, cycles);
    return (int)cudaGetLastError();
}
