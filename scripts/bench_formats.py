#!/usr/bin/env python
"""BASELINE.json's other configs as single-GPU SpMV timings, ours vs the closed library on the same buffers:
config 3 (fp32 SELL 7-pt 256^3), config 4's operator (fp64 CSR 5-pt 8192^2), COO and fp32 CSR on the R-MAT matrix."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cudalibrarysamples_b200 import cusparse_api as cs
from cudalibrarysamples_b200 import workloads as W


def timeit(op, x, y, steps=50):
    for _ in range(5):
        op(x, y, 1.0, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        op(x, y, 1.0, 0.0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps


def case(name, fmt, rows, cols, arrays, nbytes, dtype, out):
    x = W.uniform(44, cols, dtype)
    res = {}
    ys = {}
    for impl in ("b200", "cusparse"):
        api = cs.Api(impl)
        op = cs.SpMVOperator(api, fmt, rows, cols, arrays)
        y = torch.zeros(rows, dtype=dtype, device="cuda")
        us = timeit(op, x, y)
        ys[impl] = y
        res[impl] = dict(us=round(us, 2), gbs=round(nbytes / us / 1e3, 1))
        op.close()
    err = float((torch.linalg.norm(ys["b200"].double() - ys["cusparse"].double()) / torch.linalg.norm(ys["cusparse"].double())).item())
    res["rel_diff"] = err
    res["alg_MB"] = round(nbytes / 1e6, 1)
    res["speedup_vs_cusparse"] = round(res["cusparse"]["us"] / res["b200"]["us"], 3)
    print(name, json.dumps(res), flush=True)
    out[name] = res


out = {}
which = sys.argv[1:] or ["sell", "cg", "coo", "f32", "spmm"]
if "sell" in which:
    nx = 256
    off, col, val = W.laplace7_csr(nx, torch.float32)
    n = nx ** 3
    so, sc, sv = W.csr_to_sell(off, col, val, 32)
    nsl = so.numel() - 1
    case("config3_sell_f32_laplace7_256", "sell", n, n, dict(off=so, col=sc, val=sv, slice_size=32, nnz=int(col.numel())),
         W.sell_bytes(n, n, int(sv.numel()), nsl, 4), torch.float32, out)
    case("csr_f32_laplace7_256", "csr", n, n, dict(off=off, col=col, val=val), W.csr_bytes(n, n, int(col.numel()), 4), torch.float32, out)
    del off, col, val, so, sc, sv
    torch.cuda.empty_cache()
if "cg" in which:
    g = 8192
    off, col, val = W.stencil5_csr(g)
    n = g * g
    case("config4_csr_f64_stencil5_8192", "csr", n, n, dict(off=off, col=col, val=val), W.csr_bytes(n, n, int(col.numel()), 8), torch.float64, out)
    del off, col, val
    torch.cuda.empty_cache()
if "coo" in which or "f32" in which:
    rows = 1_000_000
    off, col, val = W.rmat_csr(rows)
    if "coo" in which:
        row = W.csr_to_coo_rows(off)
        case("coo_f64_rmat1m", "coo", rows, rows, dict(row=row, col=col, val=val), W.coo_bytes(rows, rows, int(col.numel()), 8), torch.float64, out)
    if "f32" in which:
        case("csr_f32_rmat1m", "csr", rows, rows, dict(off=off, col=col, val=val.float()), W.csr_bytes(rows, rows, int(col.numel()), 4), torch.float32, out)
if "spmm" in which:
    # BASELINE.json configs[4] on one GPU: fp32 CSR x dense, A 2M x 2M with 32 non-zeros per row, n = 64
    rows, per_row, n = 2_000_000, 32, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    col = torch.randint(0, rows, (rows, per_row), device="cuda", generator=g, dtype=torch.int32).sort(dim=1).values.reshape(-1).contiguous()
    off = (torch.arange(rows + 1, device="cuda", dtype=torch.int64) * per_row).to(torch.int32)
    val = W.uniform(43, rows * per_row, torch.float32)
    arrays = dict(off=off, col=col, val=val)
    B = W.uniform(46, rows * n, torch.float32)
    C0 = torch.zeros(rows * n, dtype=torch.float32, device="cuda")
    nbytes = rows * per_row * 8 + (rows + 1) * 4 + 2 * rows * n * 4          # A once, B once, C written once
    for order, oname in ((cs.CUSPARSE_ORDER_COL, "colmajor"), (cs.CUSPARSE_ORDER_ROW, "rowmajor")):
        res, cs_out = {}, {}
        for impl in ("b200", "cusparse"):
            api = cs.Api(impl)
            ts = []
            for rep in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                out_c = cs.spmm(api, rows, rows, arrays, B, C0, 1.0, 0.0, order, order, timing=(e0, e1))
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            cs_out[impl] = out_c
            us = sorted(ts[1:])[len(ts[1:]) // 2]
            res[impl] = dict(us=round(us, 1), gbs=round(nbytes / us / 1e3, 1), gflops=round(2.0 * rows * per_row * n / us / 1e3, 1))
        res["rel_diff"] = float((torch.linalg.norm(cs_out["b200"].double() - cs_out["cusparse"].double()) / torch.linalg.norm(cs_out["cusparse"].double())).item())
        res["alg_MB"] = round(nbytes / 1e6, 1)
        res["speedup_vs_cusparse"] = round(res["cusparse"]["us"] / res["b200"]["us"], 3)
        print("config5_spmm_f32_2m_n64_" + oname, json.dumps(res), flush=True)
        out["config5_spmm_f32_2m_n64_" + oname] = res
        del cs_out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_formats.json"), "w"), indent=1)
