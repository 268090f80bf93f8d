#!/bin/bash
# Round-2 call J: new paths (SpMM transposing path, opA=TRANSPOSE, short kernel + dot epilogue, COO_ALG2 forward), flat kernel
# variants (quiet-step ballot, empty rows in tail CTAs), formats, bench N=1
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2j_pytest.log 2>&1; echo "rc=$?"; tail -n 6 $OUT/r2j_pytest.log
echo "== flat3 variants"
SWEEP_SET=flat3 timeout 400 python scripts/sweep.py run rmat1m rmat10m rmat1m_f32 > $OUT/r2j_sweep_flat3.txt 2>&1; grep -E "==|us " $OUT/r2j_sweep_flat3.txt
echo "== formats"
timeout 500 python scripts/bench_formats.py > $OUT/r2j_formats.txt 2>&1; cut -c1-260 $OUT/r2j_formats.txt; cp $OUT/bench_formats.json $OUT/r2j_formats.json
echo "== bench N=1"
timeout 600 python bench.py --steps 200 --warmup 10 > $OUT/r2j_bench.json 2> $OUT/r2j_bench.err; echo "rc=$?"; tail -n 3 $OUT/r2j_bench.err
python - <<PY
import json
d = json.load(open("$OUT/r2j_bench.json"))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, d["roofline"]["kernel"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "closed", d["cusparse_same_box"].get("ms_per_step"))
print("north", d["north_star_10m"]["ours"], d["north_star_10m"]["rel_err_vs_cusparse"]); print("cg", d["cg_config4"]["value"], d["cg_config4"]["driver"][:120])
print("others", json.dumps(d["other_configs"])[:900])
PY
