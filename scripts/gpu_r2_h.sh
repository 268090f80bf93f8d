#!/bin/bash
# Round-2 call H: csr_flat_kernel with decoupled look-back (single launch): parity, variants, bench
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2h_pytest.log 2>&1; echo "rc=$?"; tail -n 6 $OUT/r2h_pytest.log
echo "== flat2 variants"
SWEEP_SET=flat2 timeout 500 python scripts/sweep.py run rmat1m rmat10m > $OUT/r2h_sweep_flat2.txt 2>&1; grep -E "==|us " $OUT/r2h_sweep_flat2.txt
echo "== formats"
timeout 300 python scripts/bench_formats.py coo f32 > $OUT/r2h_formats.txt 2>&1; cut -c1-230 $OUT/r2h_formats.txt
echo "== micro_gather (x tiles in shared memory: cost model)"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/micro_gather scripts/micro_gather.cu 2>&1 | tail -n 2
python -c "
from cudalibrarysamples_b200 import workloads as W
off,col,val=W.rmat_csr(1000000); col.cpu().numpy().tofile('/tmp/rmat_col.bin')" && timeout 200 /tmp/micro_gather /tmp/rmat_col.bin 1000000 > $OUT/r2h_micro_hot.txt 2>&1; cat $OUT/r2h_micro_hot.txt
echo "== bench N=1 (short)"
timeout 600 python bench.py --steps 200 --warmup 10 > $OUT/r2h_bench.json 2> $OUT/r2h_bench.err; echo "rc=$?"; tail -n 3 $OUT/r2h_bench.err
python - <<PY
import json
d = json.load(open("$OUT/r2h_bench.json"))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, d["roofline"]["kernel"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "closed", d["cusparse_same_box"].get("ms_per_step"))
print("north", d["north_star_10m"]["ours"], d["north_star_10m"]["rel_err_vs_cusparse"]); print("cg", d["cg_config4"]["value"])
PY
