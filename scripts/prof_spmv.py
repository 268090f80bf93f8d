#!/usr/bin/env python
"""Short driver for ncu: a few fp64 CSR SpMV launches on BASELINE.json configs[1] (R-MAT 1M x 1M, 16/row) or another
workload, with our kernels and/or the closed library.  Usage: prof_spmv.py [--impl b200|cusparse|both] [--workload ...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="both")
ap.add_argument("--workload", default="rmat1m", choices=["rmat1m", "rmat10m", "stencil5_4096", "laplace7_sell", "uniform1m"])
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()

fmt = "csr"
if a.workload == "rmat1m":
    rows = 1_000_000
    off, col, val = W.rmat_csr(rows)
    arrays = dict(off=off, col=col, val=val)
elif a.workload == "rmat10m":
    rows = 10_000_000
    off, col, val = W.rmat_csr(rows)
    arrays = dict(off=off, col=col, val=val)
elif a.workload == "stencil5_4096":
    rows = 4096 * 4096
    off, col, val = W.stencil5_csr(4096)
    arrays = dict(off=off, col=col, val=val)
elif a.workload == "uniform1m":
    rows = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    col = torch.randint(0, rows, (rows, 16), device="cuda", generator=g, dtype=torch.int32).sort(dim=1).values.reshape(-1).contiguous()
    off = (torch.arange(rows + 1, device="cuda", dtype=torch.int64) * 16).to(torch.int32)
    val = W.uniform(43, rows * 16)
    arrays = dict(off=off, col=col, val=val)
else:
    nx = 256
    rows = nx ** 3
    off, col, val = W.laplace7_csr(nx, torch.float32)
    so, sc, sv = W.csr_to_sell(off, col, val, 32)
    arrays = dict(off=so, col=sc, val=sv, slice_size=32, nnz=int(col.numel()))
    fmt = "sell"
dt = arrays["val"].dtype
x = W.uniform(44, rows, dt)
for impl in (["b200", "cusparse"] if a.impl == "both" else [a.impl]):
    api = cs.Api(impl)
    op = cs.SpMVOperator(api, fmt, rows, rows, arrays)
    y = torch.zeros(rows, dtype=dt, device="cuda")
    for _ in range(a.iters):
        op(x, y, 1.0, 0.0)
    torch.cuda.synchronize()
    op.close()
print("done", a.workload, a.impl)
