#!/bin/bash
# Standard evidence run for one round, to be executed on the GPU box:
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh rNN'
# Leaves everything under gpurun_out/<tag>_*; copy what should be judged into profiles/ afterwards.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu";  timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -n 2 $OUT/${TAG}_pytest_gpu.log
echo "== smoke";          python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
echo "== bench N=1";      python bench.py --steps 200 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print("value", d["value"], "GB/s  ms/step", d["ms_per_step"], " frac", d["roofline"]["frac"], " closed", d["cusparse_same_box"]["value"],
      " e2e", d["e2e"]["value"], " cpu", d["cpu_baseline"]["value"], d["clocks"])
PY
echo "== other configs";  timeout 400 python scripts/bench_formats.py > $OUT/${TAG}_bench_formats.txt 2>&1; cut -c1-200 $OUT/${TAG}_bench_formats.txt
echo "== CSR kernels on the standard matrices"; SWEEP_SET=none timeout 400 python scripts/sweep.py run rmat1m rmat10m uniform1m stencil5_4096 > $OUT/${TAG}_sweep_default.txt 2>&1; grep -E "==|us " $OUT/${TAG}_sweep_default.txt
echo "== ncu launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/${TAG}_bench_ncu.log 2>&1
echo "== ncu --set full (dominant kernel)"; ncu --set full --clock-control none --import-source on -k regex:"csr_tile_kernel|csr_pipe_kernel" -s 2 -c 1 -o $OUT/prof_${TAG}_csr python scripts/prof_spmv.py --impl b200 --workload rmat1m > $OUT/${TAG}_ncu.log 2>&1; tail -n 1 $OUT/${TAG}_ncu.log
echo "== every CSR kernel of the library (incl. the unselected candidates tile2 / hyb)"; SWEEP_SET=kernels timeout 400 python scripts/sweep.py run rmat1m rmat10m uniform1m stencil5_4096 > $OUT/${TAG}_sweep_kernels.txt 2>&1; grep -E "==|us " $OUT/${TAG}_sweep_kernels.txt
