#!/usr/bin/env python
"""Time cusparseSpMV (ours and the closed library) on a Matrix Market file -- e.g. a SuiteSparse matrix copied to the GPU box.
usage: python scripts/spmv_mtx.py path/to/matrix.mtx [--f32]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W
from cudalibrarysamples_b200.mtx import read_matrix_market

path = sys.argv[1]
dtype = torch.float32 if "--f32" in sys.argv else torch.float64
n, m, off, col, val = read_matrix_market(path, dtype=np.float32 if dtype == torch.float32 else np.float64)
arrays = dict(off=torch.tensor(off, device="cuda"), col=torch.tensor(col, device="cuda"), val=torch.tensor(val, device="cuda"))
x = W.uniform(44, m, dtype)
out = {"file": os.path.basename(path), "rows": n, "cols": m, "nnz": int(col.size)}
ys = {}
for impl in ("b200", "cusparse"):
    api = cs.Api(impl)
    op = cs.SpMVOperator(api, "csr", n, m, arrays)
    y = torch.zeros(n, dtype=dtype, device="cuda")
    call = op.prebuilt(x, y, 1.0, 0.0)
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10.0
    nbytes = W.csr_bytes(n, m, int(col.size), val.itemsize)
    out[impl] = {"us": round(us, 2), "gbs": round(nbytes / us / 1e3, 1)}
    if impl == "b200":
        out["kernel"] = api.last_csr_kernel()
    ys[impl] = y
    op.close()
out["rel_diff"] = float((torch.linalg.norm(ys["b200"].double() - ys["cusparse"].double()) / torch.linalg.norm(ys["cusparse"].double())).item())
print(json.dumps(out))
