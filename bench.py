#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: fp64 CSR SpMV effective HBM GB/s (and fraction of the HBM roofline).

  python bench.py [--gpus N --steps K --warmup W]          our arm: sm_100a kernels through the cuSPARSE C ABI
  python bench.py --impl reference [...]                    reference arm: the samples' host loop on the CPU cores
  torchrun --nproc-per-node N bench.py --gpus N ...         N>1: row-block shards + one NCCL all-gather of x per step

A "step" is one y = A*x (alpha=1, beta=0) over the whole matrix.  Workload at N=1 = BASELINE.json configs[1]:
R-MAT 1,000,000 x 1,000,000, 16 non-zeros/row on average, fp64 values, int32 indices.  At N>1 the
matrix grows with N (N*1M rows, "weak" scaling): every rank keeps ~16M non-zeros and receives the other ranks' x.
Timing: CUDA events on the launching stream around exactly K steps, barrier + synchronize on both sides, max over
ranks.  No L2 flush: one step streams 212 MB (> 126 MB L2) so val[]/col_ind[] cannot stay resident; x (8 MB) does.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = sys.stdout
METRIC = "csr_spmv_fp64_effective_hbm_bandwidth"
UNIT = "GB/s"
ROWS_PER_GPU = 1_000_000
AVG_NNZ = 16


def csr_bytes(rows, cols, nnz, vb=8, ib=4):
    # nnz*(val+idx) + (rows+1)*idx + cols*val (x) + rows*val (y), beta = 0
    return nnz * (vb + ib) + (rows + 1) * ib + cols * vb + rows * vb


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread for the whole run; samples inside the timed window are reported
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.ok, self._stop = [], False, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok and self.t.is_alive():
            self.t.join(1.0)

    def summary(self, t0, t1, probe=None):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        win, where = [s for s in self.samples if t0 <= s[0] <= t1], "timed region"
        if len(win) < 3 and probe is not None:
            win, where = [s for s in self.samples if probe[0] <= s[0] <= probe[1]], "same kernel looped for 1 s right after the timed region (region too short to sample)"
        if not win:
            win, where = self.samples[-5:], "last samples"
        clocks = sorted(s[1] for s in win)
        bits = 0
        for s in win:
            bits |= s[2]
        reasons = [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": self.max, "reasons": reasons,
                "samples": len(win), "window": where}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the samples' host verification loop (spmv_csr_op_example.c:307-318) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def pick_threads(off, col, val, x):
    """Host threads for the CPU arm: the fastest of {allowed CPUs, half of them (physical cores), 32}, probed with two
    repetitions each -- omp_get_max_threads() counts SMT siblings / CPUs outside the cgroup on some boxes and the loop
    then runs 20x slower."""
    from oracle import oracle as O
    try:
        allowed = len(os.sched_getaffinity(0))
    except Exception:
        allowed = O.max_threads()
    cands = sorted({max(1, min(allowed, O.max_threads())), max(1, allowed // 2), max(1, min(32, allowed))}, reverse=True)
    best_t, best = cands[0], float("inf")
    for t in cands:
        sec, _, _ = O.time_csr_f64(off, col, val, x, t, reps=2)
        if sec < best:
            best_t, best = t, sec
    return best_t


def cpu_reference(off, col, val, x, reps):
    from oracle import oracle as O
    threads = pick_threads(off, col, val, x)
    rows = off.size - 1
    best, times, _ = O.time_csr_f64(off, col, val, x, threads, reps=reps)
    gbs = csr_bytes(rows, x.size, col.size) / best / 1e9
    return gbs, threads, best, times


def run_reference_arm(args):
    """The reference's own CPU implementation of the path, all host threads, same metric / unit / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    t_gen = time.time()
    off, col, val = O.rmat_csr(ROWS_PER_GPU, avg_nnz=AVG_NNZ, seed=42, val_seed=43)
    x = O.uniform(44, ROWS_PER_GPU)
    t_gen = time.time() - t_gen
    threads = pick_threads(off, col, val, x)
    rows, nnz = off.size - 1, int(col.size)
    y = np.zeros(rows)
    for _ in range(args.warmup):
        O.time_csr_f64(off, col, val, x, threads, reps=1)
    _, times, _ = O.time_csr_f64(off, col, val, x, threads, reps=args.steps)
    total = float(sum(times))
    ms = 1e3 * total / args.steps
    gbs = csr_bytes(rows, rows, nnz) / (total / args.steps) / 1e9
    sample = (f"one full y=A*x per step on the 1,000,000-row R-MAT shard (nnz={nnz}); OpenMP dynamic row chunks, "
              f"{threads} threads; matrix generated on the CPU in {t_gen:.0f} s (not timed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, rows, nnz) | {"reference_arm": "CPU host loop (spmv_csr_op_example.c:307-318 restated, oracle/spmv_oracle.c); per-rank shard size"},
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops": round(2 * nnz / (total / args.steps) / 1e9, 3),
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def workload_config(n_gpus, rows, nnz):
    return {
        "workload": f"fp64 CSR SpMV y=A*x, synthetic R-MAT {rows}x{rows}, nnz={nnz} (avg {nnz / max(rows, 1):.2f}/row), "
                    "(a,b,c,d)=(0.57,0.19,0.19,0.05), seed 42, duplicates merged, columns sorted, val,x~U(-1,1), int32 indices",
        "baseline_config": "BASELINE.json configs[1] (R-MAT 1M x 1M avg 16 nnz/row, single B200); N>1 grows the matrix to N*1M rows",
        "rows": rows, "cols": rows, "nnz": nnz, "alpha": 1.0, "beta": 0.0,
        "parallelism": "single GPU" if n_gpus == 1 else f"{n_gpus} row-block shards of A and y (nnz-balanced), x in equal blocks, one NCCL all-gather of x per step",
        "l2": "no flush: 212 MB streamed per step per GPU > 126 MB L2; x stays L2-resident by design",
    }


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def prebuilt_spmv_call(cs, op, x, y):
    """One cusparseSpMV through the C ABI with the ctypes arguments built once (keeps host overhead ~2 us/step)."""
    api = op.api
    api.cusparseDnVecSetValues(op.vecX, x)
    api.cusparseDnVecSetValues(op.vecY, y)
    one, zero = C.c_double(1.0), C.c_double(0.0)
    fn = api.lib.cusparseSpMV
    argv = (op.handle, C.c_int(cs.CUSPARSE_OPERATION_NON_TRANSPOSE), C.cast(C.pointer(one), C.c_void_p), op.mat, op.vecX,
            C.cast(C.pointer(zero), C.c_void_p), op.vecY, C.c_int(cs.CUDA_R_64F), C.c_int(op.alg),
            C.c_void_p(op.buffer.data_ptr()))

    def call(_keep=(one, zero)):
        st = fn(*argv)
        if st != 0:
            raise RuntimeError(f"cusparseSpMV status {st}")
    return call


def time_steps(torch, fn, steps, dist_on):
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, t0, t1


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from cudalibrarysamples_b200 import cusparse_api as cs
    from cudalibrarysamples_b200 import workloads as W
    from cudalibrarysamples_b200.sharded import ShardedCsr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dist_on = world > 1
    if dist_on:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; the contract is ONE JSON line on stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        # Plain stream order for the fix-up launch next to NCCL kernels: the programmatic early launch buys ~0.5 us on
        # 100 us, and one 4-GPU run of this script hung for reasons we could not reproduce (profiles/README.md).
        os.environ.setdefault("B200SPMV_NO_PDL", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    sampler.start()
    api = cs.Api("b200")
    peak, peak_src = measured_peak()

    rows = ROWS_PER_GPU * world
    off, col, val = W.rmat_csr(rows, avg_nnz=AVG_NNZ, seed=42, val_seed=43)
    nnz = int(col.numel())
    x = W.uniform(44, rows)
    total_bytes = csr_bytes(rows, rows, nnz)
    launches_per_step = 2 if nnz >= 12 * rows else 1   # csr_tile_kernel + csr_fixup_kernel, or the single persistent csr_pipe_kernel

    if not dist_on:
        op = cs.SpMVOperator(api, "csr", rows, rows, dict(off=off, col=col, val=val), preprocess=True)
        y = torch.zeros(rows, dtype=torch.float64, device="cuda")
        step = prebuilt_spmv_call(cs, op, x, y)
        kernel_bytes = total_bytes
        local = dict(rows=rows, nnz=nnz)
    else:
        def make_local(r, c, arrays):
            lop = cs.SpMVOperator(api, "csr", r, c, arrays, preprocess=True)
            make_local.op = lop
            return lop
        sh = ShardedCsr(off, col, val, rank, world, make_local, exchange="allgather")   # R-MAT rows read every x block
        del off, col, val
        torch.cuda.empty_cache()
        xs = sh.new_x_shard(x)
        ys = sh.new_y_shard()
        lop = make_local.op
        local_call = prebuilt_spmv_call(cs, lop, sh.x_full, ys)

        def step():
            dist.all_gather_into_tensor(sh.x_full, xs)
            local_call()
        kernel_bytes = csr_bytes(sh.rows, sh.cols_padded, sh.nnz)
        local = dict(rows=sh.rows, nnz=sh.nnz, x_block=sh.x_block)
        # parity of the sharded product at full size: every rank checks its y shard against the closed library run
        # on the same local arrays and the gathered x (max relative difference over ranks goes into the JSON line)
        step()
        torch.cuda.synchronize()
        capi = cs.Api("cusparse")
        cop = cs.SpMVOperator(capi, "csr", sh.rows, sh.cols_padded, dict(off=sh.off, col=sh.col, val=sh.val), preprocess=True)
        yc = torch.zeros_like(ys)
        cop(sh.x_full, yc, 1.0, 0.0)
        torch.cuda.synchronize()
        shard_rel = (torch.linalg.norm(ys - yc) / torch.linalg.norm(yc)).reshape(1)
        dist.all_reduce(shard_rel, op=dist.ReduceOp.MAX)
        local["max_rel_diff_vs_cusparse_over_ranks"] = float(shard_rel.item())
        assert local["max_rel_diff_vs_cusparse_over_ranks"] < 1e-12
        cop.close()

    for _ in range(max(args.warmup, 3)):
        step()
    ms_total, t0, t1 = time_steps(torch, step, args.steps, dist_on)
    ms_step = ms_total / args.steps
    value = total_bytes / (ms_step * 1e-3) / 1e9

    # dominant kernel alone (no collective in the loop): average launch duration over the same K launches
    if dist_on:
        for _ in range(3):
            local_call()
        ms_k, _, _ = time_steps(torch, local_call, args.steps, False)
        kern_ms = ms_k / args.steps
    else:
        kern_ms = ms_step
    achieved = kernel_bytes / (kern_ms * 1e-3) / 1e9

    # clocks: if the timed region was too short for NVML's sampling period, loop the same step for ~1 s and sample that
    probe = None
    if (t1 - t0) < 0.25:
        p0 = time.perf_counter()
        while time.perf_counter() - p0 < 1.0:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        probe = (p0, time.perf_counter())
    clocks = sampler.summary(t0, t1, probe)
    sampler.stop()

    # ---- e2e: the user-facing call with HOST buffers; x goes up and y comes back every step (A stays resident, as in
    #      cg_example.c:327-362 where the matrix is uploaded once and only vectors move) ----
    e2e = None
    if not dist_on:
        hx = x.cpu().pin_memory()
        hy = torch.empty(rows, dtype=torch.float64).pin_memory()
        dx, dy = torch.empty_like(x), torch.empty_like(y)

        def e2e_step():
            dx.copy_(hx, non_blocking=True)
            op(dx, dy, 1.0, 0.0)
            hy.copy_(dy, non_blocking=True)
        for _ in range(3):
            e2e_step()
        ms_e, _, _ = time_steps(torch, e2e_step, args.steps, False)
        e2e = {"value": round(total_bytes / (ms_e / args.steps * 1e-3) / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": rows * 8, "d2h_bytes_per_step": rows * 8, "ms_per_step": round(ms_e / args.steps, 4),
               "what": "pinned host x -> device, cusparseSpMV through the C ABI, device y -> pinned host, every step; A resident"}
        assert torch.equal(hy, y.cpu()), "e2e result differs from the device-resident result"
    else:
        hx = xs.cpu().pin_memory()
        hy = torch.empty(max(sh.rows, 1), dtype=torch.float64).pin_memory()[:sh.rows]

        def e2e_step():
            xs.copy_(hx, non_blocking=True)
            step()
            hy.copy_(ys, non_blocking=True)
        for _ in range(3):
            e2e_step()
        ms_e, _, _ = time_steps(torch, e2e_step, args.steps, True)
        e2e = {"value": round(total_bytes / (ms_e / args.steps * 1e-3) / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": sh.x_block * 8 * world, "d2h_bytes_per_step": rows * 8,
               "ms_per_step": round(ms_e / args.steps, 4),
               "what": "per rank: pinned host x shard -> device, all-gather + cusparseSpMV, device y shard -> pinned host"}

    # ---- closed cusparseSpMV (csrmv_v3_kernel, sm_100 SASS) on the same device buffers: the on-box bar to beat ----
    closed = None
    if not dist_on and not args.no_cusparse:
        try:
            capi = cs.Api("cusparse")
            cop = cs.SpMVOperator(capi, "csr", rows, rows, dict(off=off, col=col, val=val), preprocess=True)
            y2 = torch.zeros_like(y)
            cstep = prebuilt_spmv_call(cs, cop, x, y2)
            for _ in range(max(args.warmup, 3)):
                cstep()
            ms_c, _, _ = time_steps(torch, cstep, args.steps, False)
            rel = float((torch.linalg.norm(y - y2) / torch.linalg.norm(y2)).item())
            closed = {"value": round(total_bytes / (ms_c / args.steps * 1e-3) / 1e9, 3), "unit": UNIT,
                      "ms_per_step": round(ms_c / args.steps, 4), "frac_of_peak": round(total_bytes / (ms_c / args.steps * 1e-3) / 1e9 / peak, 4),
                      "rel_diff_vs_ours": rel, "what": "closed libcusparse.so.12 cusparseSpMV (the copy torch ships; preprocessed), same buffers, same loop"}
            cop.close()
        except Exception as e:  # pragma: no cover
            closed = {"error": repr(e)}

    # ---- cpu_baseline: the oracle port on this box's host cores, rank 0, N=1 only, bounded sample ----
    cpu = None
    if not dist_on and not args.no_cpu:
        h_off, h_col, h_val, h_x = (t.cpu().numpy() for t in (off, col, val, x))
        gbs, threads, best, times = cpu_reference(h_off, h_col, h_val, h_x, reps=5)
        # correctness of the measured result against the oracle, at full size
        from oracle import oracle as O
        ref = O.spmv_csr(h_off, h_col, h_val, h_x, threads=threads)
        rel = float(np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref))
        assert rel < 1e-12, f"GPU result differs from the CPU oracle: {rel}"
        cpu = {"value": round(gbs, 3), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the full {rows}-row matrix, 5 repetitions, min ({best * 1e3:.1f} ms); OpenMP dynamic row chunks",
               "gpu_vs_oracle_rel_err": rel}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, rows, nnz),
            "gflops": round(2 * nnz / (ms_step * 1e-3) / 1e9, 3),
            "frac_of_hbm_peak": round(value / (peak * world), 4),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "kernel": "b200::csr_tile_kernel<double>", "kernel_avg_us": round(kern_ms * 1e3, 3),
                         "algorithmic_bytes_per_launch": kernel_bytes, "peak_source": peak_src,
                         "per_rank": local},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "cusparse_same_box": closed, "impl": "b200",
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                line["roofline"]["traffic"] = json.load(open(prof)).get("csr_tile_kernel_f64_rmat1m_dram_bytes")
            except Exception:
                pass
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun hints) also write to fd 1, so
    # everything but our line is re-routed to stderr.
    global _REAL_STDOUT
    # Watchdog: a distributed hang must not wedge the GPU box -- dump every thread's stack and exit after the limit.
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900" if "--impl" in sys.argv and "reference" in sys.argv else "240")), exit=True)
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cusparse", action="store_true", help="skip the closed-library comparison leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
