#!/usr/bin/env python
"""A few CG iterations of BASELINE.json configs[3] (5-pt Poisson 8192^2, fp64) on one GPU with plain stream launches, for an
ncu launch list: which share of an iteration is the SpMV (+ dot epilogue) and which the fused BLAS-1 kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W
from cudalibrarysamples_b200.cg import FusedCgSolver
from cudalibrarysamples_b200.sharded import ShardedCsr

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
api = cs.Api("b200")
off, col, val = W.stencil5_csr(grid)
n = grid * grid
sh = ShardedCsr(off, col, val, 0, 1, lambda r, c, a: cs.SpMVOperator(api, "csr", r, c, a, preprocess=True), balance="rows")
del off, col, val
b = sh.new_y_shard()
sh.spmv(sh.new_x_shard(torch.ones(n, dtype=torch.float64, device="cuda")), b, alpha=0.75, beta=0.0)
solver = FusedCgSolver(sh, b, use_graph=False)
x, norms = solver.run(6)
torch.cuda.synchronize()
print("done", norms, solver.describe()[:80])
