#!/bin/bash
# Round-2 call K: which GPU tests are slow; CG iteration breakdown; ncu --set full of the new kernels
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu --durations"
timeout 900 python -m pytest tests -m gpu -x -q --durations=40 > $OUT/r2k_pytest.log 2>&1; echo "rc=$?"; tail -n 50 $OUT/r2k_pytest.log
echo "== CG iteration: ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_short|update_xr|update_p|dot_kernel" -c 40 --csv --log-file $OUT/r2k_launches_cg.csv python scripts/prof_cg.py > $OUT/r2k_prof_cg.log 2>&1; echo "rc=$?"; tail -n 2 $OUT/r2k_prof_cg.log
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
echo "== ncu --set full: csr_short_kernel (fp64 5-pt 4096^2), spmm kernels, transpose kernel"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"csr_short_kernel|spmm_csr|spmm_transpose|csr_pipe_kernel" -c 10 -o $OUT/prof_r2k_new python scripts/prof_all.py > $OUT/r2k_ncu_new.log 2>&1; echo "rc=$?"; tail -n 2 $OUT/r2k_ncu_new.log
