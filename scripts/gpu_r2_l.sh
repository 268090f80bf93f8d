#!/bin/bash
# Round-2 call L: the round's final evidence run on one B200 (everything lands in gpurun_out/r2l_*; judged copies go to profiles/)
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q --durations=30 > $OUT/r2l_pytest.log 2>&1; echo "rc=$?"; tail -n 40 $OUT/r2l_pytest.log | cut -c1-150
echo "== bench N=1"
timeout 600 python bench.py --steps 200 --warmup 10 > $OUT/r2l_bench.json 2> $OUT/r2l_bench.err; echo "rc=$?"; tail -n 3 $OUT/r2l_bench.err
python - <<PY
import json
d = json.load(open("$OUT/r2l_bench.json"))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, d["roofline"]["kernel"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "closed", d["cusparse_same_box"].get("ms_per_step"), "toolkit", d["cusparse_toolkit"].get("us_per_spmv"), d["clocks"])
print("cpu", d["cpu_baseline"]["value"]); print("north", d["north_star_10m"]["ours"]); print("cg", d["cg_config4"]["value"])
PY
echo "== reference arm"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/r2l_bench_reference.json 2>> $OUT/r2l_bench.err; echo "rc=$?"; cut -c1-200 $OUT/r2l_bench_reference.json
echo "== formats"
timeout 400 python scripts/bench_formats.py > $OUT/r2l_formats.txt 2>&1; cut -c1-230 $OUT/r2l_formats.txt; cp $OUT/bench_formats.json $OUT/r2l_formats.json
echo "== ncu launch list of the bench command"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2l_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-cusparse --no-extra > $OUT/r2l_bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu --set full, main kernel of every format"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_flat_kernel|csr_short_kernel|coo_seg_kernel|sell32_kernel|spmm_csr|spmm_transpose" -c 16 -o $OUT/prof_r2l_all python scripts/prof_all.py > $OUT/r2l_ncu_all.log 2>&1; echo "rc=$?"; tail -n 2 $OUT/r2l_ncu_all.log
