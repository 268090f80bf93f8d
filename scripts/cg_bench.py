#!/usr/bin/env python
"""BASELINE.json config 4: fp64 CG on the 5-pt Poisson matrix of cuSPARSE/cg/cg_example.c:71-128 scaled to grid^2 rows,
200 fixed iterations, row-sharded over N GPUs; reports iterations/s (max time over ranks, CUDA events) for our SpMV and
for the closed library's SpMV inside the same loop.

  python scripts/cg_bench.py [--grid 8192 --iters 200]                                  (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... scripts/cg_bench.py --gpus N   (N GPUs)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W
from cudalibrarysamples_b200.cg import conjugate_gradient
from cudalibrarysamples_b200.sharded import ShardedCsr

sys.stdout.flush()
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)   # only our JSON line goes to the real stdout
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--grid", type=int, default=8192)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
assert world == a.gpus
torch.cuda.set_device(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = a.grid * a.grid
assert n % world == 0
off, col, val = W.stencil5_csr(a.grid)
out = {}
for impl in ("b200", "cusparse"):
    api = cs.Api(impl)

    def make_local(r, c, arrays, api=api):
        return cs.SpMVOperator(api, "csr", r, c, arrays, preprocess=True)

    sh = ShardedCsr(off, col, val, rank, world, make_local, balance="rows")
    exch = sh.exchange + (f" ({sh.exchanged_elements * 8} B per step over all ranks)" if sh.exchange == "halo" else "")
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    b = sh.new_y_shard()
    sh.spmv(sh.new_x_shard(ones), b, alpha=0.75, beta=0.0)      # b = 0.75 * A * 1 (cg_example.c:405-418)
    del ones
    conjugate_gradient(sh, b, 3)                                # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x, norms = conjugate_gradient(sh, b, a.iters)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # true residual ||b - A x|| with one more product (cg_example.c:289-300)
    r = b.clone()
    sh.spmv(x, r, alpha=-1.0, beta=1.0)
    rr = torch.dot(r, r).reshape(1)
    if world > 1:
        dist.all_reduce(rr)
    out[impl] = dict(iters_per_s=round(a.iters / (float(ms.item()) * 1e-3), 2), ms_per_iter=round(float(ms.item()) / a.iters, 4),
                     r0=float(norms[0].item()), r_end=float(norms[-1].item()), true_residual=float(rr.sqrt().item()))
    sh.local_op.close()
    del sh, b, x, r
    torch.cuda.empty_cache()
if rank == 0:
    line = {"metric": "cg_iterations_per_second", "config": f"BASELINE.json configs[3]: fp64 CG, 5-pt Poisson {a.grid}^2 ({n} rows), "
            f"{a.iters} fixed iterations, unpreconditioned, row-sharded over {world} GPU(s), one all-gather of p per iteration",
            "n_gpus": world, "ours": out["b200"], "closed_library_spmv_same_loop": out["cusparse"],
            "speedup": round(out["b200"]["iters_per_s"] / out["cusparse"]["iters_per_s"], 3)}
    line["exchange"] = exch
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
