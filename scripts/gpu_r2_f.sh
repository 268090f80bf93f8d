#!/bin/bash
# Round-2 call F: full GPU test-suite, flat variants (steps per chunk, nzrow look-ahead), bench N=1
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu (everything)"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2f_pytest.log 2>&1; echo "rc=$?"; tail -n 8 $OUT/r2f_pytest.log
echo "== flat variants"
SWEEP_SET=flat timeout 500 python scripts/sweep.py run rmat1m rmat10m > $OUT/r2f_sweep_flat.txt 2>&1; grep -E "==|us " $OUT/r2f_sweep_flat.txt
echo "== coo / f32 / sell / cg operator"
timeout 400 python scripts/bench_formats.py > $OUT/r2f_formats.txt 2>&1; cut -c1-220 $OUT/r2f_formats.txt
echo "== bench N=1"
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err; echo "rc=$?"; tail -n 5 $OUT/r2f_bench.err
python - <<PY
import json
d = json.load(open("$OUT/r2f_bench.json"))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, d["roofline"]["kernel"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "toolkit", d["cusparse_toolkit"].get("us_per_spmv"))
print("north", d["north_star_10m"]); print("cg", d["cg_config4"])
PY
