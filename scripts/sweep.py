#!/usr/bin/env python
"""Parameter sweeps for the CSR tile kernel.

  python scripts/sweep.py build            (CPU container)  nvcc-build one .so per variant into build/variants/
  python scripts/sweep.py run [workloads]  (GPU box)        time every variant + the closed library on each workload

Variants are -D overrides of the tunables at the top of csrc/spmv_csr.cu.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "cudalibrarysamples_b200", "build", "variants")

VARIANTS = {
    # tag: dict(TILE, LONG, BLOCK, BATCH, MIN_CTAS)
    "t2048_b256_k4": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=1),
    "t2048_b256_k4_occ6": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=6),
    "t2048_b256_k4_occ8": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=8),
    "t2048_b256_k4_occ5": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=5),
    "t2048_b256_k2_occ8": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=2, MIN=8),
    "t2048_b256_k8": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=8, MIN=1),
    "t2048_b128_k4": dict(TILE=2048, LONG=512, BLOCK=128, BATCH=4, MIN=1),
    "t2048_b128_k8": dict(TILE=2048, LONG=512, BLOCK=128, BATCH=8, MIN=1),
    "t1024_b128_k4": dict(TILE=1024, LONG=256, BLOCK=128, BATCH=4, MIN=1),
    "t1024_b256_k4": dict(TILE=1024, LONG=256, BLOCK=256, BATCH=4, MIN=1),
    "t4096_b256_k8": dict(TILE=4096, LONG=1024, BLOCK=256, BATCH=8, MIN=1),
    "t4096_b512_k4": dict(TILE=4096, LONG=1024, BLOCK=512, BATCH=4, MIN=1),
    "t3072_b256_k4": dict(TILE=3072, LONG=512, BLOCK=256, BATCH=4, MIN=1),
}


ABL = {f"abl{a}_occ6": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=6, ABL=a) for a in (1, 2, 3)}
if os.environ.get("SWEEP_SET") == "occ":
    VARIANTS = {f"t2048_b256_k{k}_occ{o}": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=k, MIN=o) for k in (4, 8) for o in (4, 5, 6, 7)}
    VARIANTS.update({f"t1024_b128_k{k}_occ{o}": dict(TILE=1024, LONG=256, BLOCK=128, BATCH=k, MIN=o) for k in (4, 8) for o in (8, 12)})
    VARIANTS.update({f"t1024_b256_k4_occ{o}": dict(TILE=1024, LONG=256, BLOCK=256, BATCH=4, MIN=o) for o in (6, 8)})
if os.environ.get("SWEEP_SET") == "pipe":
    VARIANTS = {f"pipe_t2048_b256_o{po}_occ{o}": dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=o, KERNEL=1, OFFS=po)
                for po in (2, 4) for o in (3, 4, 5, 6)}
    VARIANTS.update({f"pipe_t1024_b128_o{po}_occ{o}": dict(TILE=1024, LONG=256, BLOCK=128, BATCH=4, MIN=o, KERNEL=1, OFFS=po)
                     for po in (4,) for o in (6, 8, 10, 12)})
    VARIANTS.update({f"pipe_t1024_b256_o2_occ{o}": dict(TILE=1024, LONG=256, BLOCK=256, BATCH=4, MIN=o, KERNEL=1, OFFS=2)
                     for o in (4, 6, 8)})
    VARIANTS.update({f"tile_t1024_b128_k8_occ12": dict(TILE=1024, LONG=256, BLOCK=128, BATCH=8, MIN=12, KERNEL=0)})
if os.environ.get("SWEEP_SET") == "red":
    VARIANTS = {}
    for rr, ru in ((4, 2), (2, 4), (4, 1), (2, 2)):
        for o in (3, 4, 5):
            VARIANTS[f"pipe_r{rr}u{ru}_occ{o}"] = dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=o, KERNEL=1, OFFS=4, RR=rr, RU=ru)
        VARIANTS[f"tile_r{rr}u{ru}_occ5"] = dict(TILE=2048, LONG=512, BLOCK=256, BATCH=4, MIN=5, KERNEL=0, RR=rr, RU=ru)
        VARIANTS[f"pipe1024_r{rr}u{ru}_occ8"] = dict(TILE=1024, LONG=256, BLOCK=128, BATCH=4, MIN=8, KERNEL=1, OFFS=4, RR=rr, RU=ru)
if os.environ.get("SWEEP_SET") == "ablate":
    VARIANTS = dict(ABL, t2048_b256_k4_occ6=VARIANTS["t2048_b256_k4_occ6"])


def flags(v):
    if "ABL" in v:
        return flags({k: x for k, x in v.items() if k != "ABL"}) + [f"-DB200_CSR_ABLATE={v['ABL']}"]
    if "RR" in v:
        return flags({k: x for k, x in v.items() if k not in ("RR", "RU")}) + [f"-DB200_CSR_RED_ROWS={v['RR']}", f"-DB200_CSR_RED_U={v['RU']}"]
    if "KERNEL" in v:
        extra = [f"-DB200_CSR_KERNEL={v['KERNEL']}"] + ([f"-DB200_CSR_PIPE_OFFS={v['OFFS']}"] if "OFFS" in v else [])
        return flags({k: x for k, x in v.items() if k not in ("KERNEL", "OFFS")}) + extra
    return [f"-DB200_CSR_TILE_ITEMS={v['TILE']}", f"-DB200_CSR_LONG_ROW={v['LONG']}", f"-DB200_CSR_BLOCK={v['BLOCK']}",
            f"-DB200_CSR_BATCH={v['BATCH']}", f"-DB200_CSR_MIN_CTAS={v['MIN']}"]


def build():
    from cudalibrarysamples_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for tag, v in VARIANTS.items():
        out = os.path.join(VDIR, f"libb200spmv_{tag}.so")
        b.build_native(extra_flags=flags(v), out_path=out, tag="v_" + tag)
        log = open(os.path.join(ROOT, "cudalibrarysamples_b200", "build", "v_" + tag, "build.log")).read()
        i = max(log.find("csr_pipe_kernelIdEE"), log.find("csr_tile_kernelIdEE")) if "-DB200_CSR_KERNEL=0" not in " ".join(flags(v)) else log.find("csr_tile_kernelIdEE")
        regs = log[i:i + 400].split("Used ")[1].split(",")[0] if i >= 0 else "?"
        print(tag, regs)


def make_workload(name):
    import torch
    from cudalibrarysamples_b200 import workloads as W
    if name.startswith("rmat"):
        rows = {"rmat1m": 1_000_000, "rmat10m": 10_000_000, "rmat4m": 4_000_000}[name]
        off, col, val = W.rmat_csr(rows)
    elif name == "uniform1m":
        rows = 1_000_000
        g = torch.Generator(device="cuda").manual_seed(1)
        col = torch.randint(0, rows, (rows, 16), device="cuda", generator=g, dtype=torch.int32).sort(dim=1).values.reshape(-1).contiguous()
        off = (torch.arange(rows + 1, device="cuda", dtype=torch.int64) * 16).to(torch.int32)
        val = W.uniform(43, rows * 16)
    elif name.startswith("stencil5_"):
        g = int(name.split("_")[1])
        rows = g * g
        off, col, val = W.stencil5_csr(g)
    elif name.startswith("laplace7_"):
        nx = int(name.split("_")[1])
        rows = nx ** 3
        off, col, val = W.laplace7_csr(nx)
    else:
        raise ValueError(name)
    return rows, off, col, val


def run(workloads, variants=None, steps=100):
    import torch
    from cudalibrarysamples_b200 import cusparse_api as cs
    from cudalibrarysamples_b200 import workloads as W
    results = {}
    libs = [("default", None)] + [(t, os.path.join(VDIR, f"libb200spmv_{t}.so")) for t in VARIANTS if (variants is None or t in variants)]
    libs = [(t, p) for t, p in libs if p is None or os.path.exists(p)]
    for wl in workloads:
        rows, off, col, val = make_workload(wl)
        nnz = int(col.numel())
        x = W.uniform(44, rows)
        nbytes = W.csr_bytes(rows, rows, nnz, 8)
        ref = None
        print(f"== {wl}: rows={rows} nnz={nnz} alg_bytes={nbytes / 1e6:.1f} MB", flush=True)
        for tag, path in libs + [("cusparse", "closed")]:
            api = cs.Api("cusparse") if tag == "cusparse" else cs.Api("b200", lib_path=path)
            op = cs.SpMVOperator(api, "csr", rows, rows, dict(off=off, col=col, val=val))
            y = torch.zeros(rows, dtype=torch.float64, device="cuda")
            for _ in range(5):
                op(x, y, 1.0, 0.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                op(x, y, 1.0, 0.0)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / steps
            if ref is None:
                ref = y.clone()
            err = float((torch.linalg.norm(y - ref) / torch.linalg.norm(ref)).item())
            print(f"  {tag:24s} {us:9.2f} us  {nbytes / us / 1e3:8.1f} GB/s  relerr_vs_first {err:.1e}", flush=True)
            results.setdefault(wl, {})[tag] = dict(us=us, gbs=nbytes / us / 1e3, err=err)
            op.close()
        del off, col, val, x
        torch.cuda.empty_cache()
    return results


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        wls = sys.argv[2:] or ["rmat1m", "uniform1m", "stencil5_4096"]
        res = run(wls)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
