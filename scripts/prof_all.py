#!/usr/bin/env python
"""Driver for ONE ncu --set full pass over the main kernel of every format / type the library serves: a few launches each of
fp64 CSR R-MAT 1M (csr_flat_kernel), fp32 CSR R-MAT 1M, fp64 COO R-MAT 1M (coo_seg_kernel), fp32 SELL 7-pt 256^3
(sell32_kernel, BASELINE config 3), fp64 CSR 5-pt 4096^2 (regular rows) and the SpMM kernel (config 5 at 1/8 size)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W

api = cs.Api(sys.argv[1] if len(sys.argv) > 1 else "b200")
ITERS = 2


def run(fmt, rows, arrays, dt):
    x = W.uniform(44, rows, dt)
    y = torch.zeros(rows, dtype=dt, device="cuda")
    op = cs.SpMVOperator(api, fmt, rows, rows, arrays)
    for _ in range(ITERS):
        op(x, y, 1.0, 0.0)
    torch.cuda.synchronize()
    op.close()


rows = 1_000_000
off, col, val = W.rmat_csr(rows)
run("csr", rows, dict(off=off, col=col, val=val), torch.float64)
run("csr", rows, dict(off=off, col=col, val=val.float()), torch.float32)
run("coo", rows, dict(row=W.csr_to_coo_rows(off), col=col, val=val), torch.float64)
del off, col, val
nx = 256
off, col, val = W.laplace7_csr(nx, torch.float32)
so, sc, sv = W.csr_to_sell(off, col, val, 32)
run("sell", nx ** 3, dict(off=so, col=sc, val=sv, slice_size=32, nnz=int(col.numel())), torch.float32)
del off, col, val, so, sc, sv
off, col, val = W.stencil5_csr(4096)
run("csr", 4096 * 4096, dict(off=off, col=col, val=val), torch.float64)
del off, col, val
torch.cuda.empty_cache()
if True:
    try:
        r = 250_000
        g = torch.Generator(device="cuda").manual_seed(1)
        c = torch.randint(0, r, (r, 32), device="cuda", generator=g, dtype=torch.int32).sort(dim=1).values.reshape(-1).contiguous()
        o = (torch.arange(r + 1, device="cuda", dtype=torch.int64) * 32).to(torch.int32)
        v = W.uniform(43, r * 32, torch.float32)
        B = W.uniform(45, r * 64, torch.float32)                        # column-major r x 64, tight leading dimension
        C0 = torch.zeros(r * 64, dtype=torch.float32, device="cuda")
        for _ in range(ITERS):
            cs.spmm(api, r, r, dict(off=o, col=c, val=v), B, C0)
        torch.cuda.synchronize()
    except Exception as e:
        print("spmm leg skipped:", repr(e))
print("done")
