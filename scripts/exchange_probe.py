#!/usr/bin/env python
"""N-GPU probe of the pieces of the x exchange (run under torchrun): device-side barrier, copy-engine peer copies, SM pull
kernel -- each timed alone with CUDA events, max over ranks.  Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

from cudalibrarysamples_b200 import lib as _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = _lib.shim()
blk = 1_000_000
sym = symm.empty(blk, dtype=torch.float64, device="cuda"); sym.fill_(rank + 1.0)
hdl = symm.rendezvous(sym, group=dist.group.WORLD)
peer = [hdl.get_buffer(r, (blk,), torch.float64) for r in range(world)]
flags = symm.empty(world, dtype=torch.int64, device="cuda"); flags.zero_(); torch.cuda.synchronize()
fh = symm.rendezvous(flags, group=dist.group.WORLD)
ptrs = torch.tensor([fh.get_buffer(r, (world,), torch.int64).data_ptr() for r in range(world)], dtype=torch.int64, device="cuda")
epoch = torch.zeros(1, dtype=torch.int64, device="cuda")
x_full = torch.zeros(world * blk, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); dist.barrier()
main = torch.cuda.current_stream()
streams = [torch.cuda.Stream() for _ in range(4)]
order = [(rank + 1 + i) % world for i in range(world - 1)]


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 2)


def barrier():
    rc = L.b200peer_barrier(C.c_void_p(main.cuda_stream), C.c_void_p(ptrs.data_ptr()), C.c_void_p(epoch.data_ptr()), C.c_int(rank), C.c_int(world), C.c_double(30.0))
    assert rc == 0


def ce_copy(chunks, nstreams):
    def fn():
        ev = torch.cuda.Event(); ev.record(main)
        k = 0
        piece = (blk + chunks - 1) // chunks
        for h in order:
            for c in range(chunks):
                lo, hi = c * piece, min((c + 1) * piece, blk)
                s = streams[k % nstreams]
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    x_full[h * blk + lo:h * blk + hi].copy_(peer[h][lo:hi], non_blocking=True)
                k += 1
        for s in streams[:nstreams]:
            main.wait_stream(s)
    return fn


def sm_pull(ctas):
    def fn():
        for h in order:
            rc = L.b200peer_pull(C.c_void_p(main.cuda_stream), C.c_void_p(x_full.data_ptr() + h * blk * 8), C.c_void_p(peer[h].data_ptr()), C.c_size_t(blk * 8), C.c_int(ctas))
            assert rc == 0
    return fn


out = {"world": world, "bytes_per_peer": blk * 8}
out["barrier_us"] = timed(barrier)
out["local_d2d_copy_us"] = timed(lambda: x_full[rank * blk:(rank + 1) * blk].copy_(sym))
for ch, ns in [(1, 1), (4, 4), (1, 4)]:
    out[f"copy_engine_{ch}chunks_{ns}streams_us"] = timed(ce_copy(ch, ns))
for ctas in [8, 16, 32, 64, 148, 296]:
    out[f"sm_pull_{ctas}ctas_us"] = timed(sm_pull(ctas))
sm_pull(64)(); torch.cuda.synchronize()
ok = all(bool((x_full[h * blk:(h + 1) * blk] == h + 1.0).all()) for h in order)
out["pulled_data_correct"] = ok
out["nccl_all_gather_us"] = timed(lambda: dist.all_gather_into_tensor(x_full, sym))
if rank == 0:
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
