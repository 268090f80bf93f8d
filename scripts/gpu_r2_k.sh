#!/bin/bash
# Round-2 call K: the paths added since call J (8-pass CG kernels), slowest GPU tests, flat kernel load-batch variants, CG breakdown
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest: new paths + durations of the heavy suspects"
timeout 600 python -m pytest tests -m gpu -x -q --durations=25 -k "fused_cg or dot_product or short_kernel_is_chosen or transposed or alg2 or north_star or config5 or edge_profiles and short" > $OUT/r2k_pytest.log 2>&1; echo "rc=$?"; tail -n 36 $OUT/r2k_pytest.log
echo "== flat4 variants (load batch depth / resident CTAs)"
SWEEP_SET=flat4 timeout 300 python scripts/sweep.py run rmat1m rmat1m_f32 rmat10m > $OUT/r2k_sweep_flat4.txt 2>&1; grep -E "==|us " $OUT/r2k_sweep_flat4.txt
echo "== CG iteration: ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_short|update_|dot_kernel" -c 40 --csv --log-file $OUT/r2k_launches_cg.csv python scripts/prof_cg.py > $OUT/r2k_prof_cg.log 2>&1; echo "rc=$?"; tail -n 2 $OUT/r2k_prof_cg.log
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/r2k_launches_cg.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i; break
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) > mv:
        v = float(r[mv].replace(",", "")); v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        d[r[kn].split("(")[0][:60]].append(v)
for k, v in d.items(): print(f"{k:62s} n={len(v):3d} mean={sum(v)/len(v):9.1f} us")
PY
echo "== CG bench leg alone"
python - <<PY
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from cudalibrarysamples_b200 import cusparse_api as cs, workloads as W
import torch.distributed as dist
out = bench.cg_leg(torch, dist, cs, W, cs.Api("b200"), 0, 1)
print(out["value"], out["ms_per_iteration"], out["driver"][:160])
PY
