#!/bin/bash
# Round-2 call E: flat CSR plan + csr_flat_kernel
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (flat / kernels)"
timeout 600 python -m pytest tests -m gpu -x -q -k "flat or every_csr_kernel or coo or plan or spmm or matrix_market" > $OUT/r2e_pytest.log 2>&1; echo "rc=$?"; tail -n 12 $OUT/r2e_pytest.log
echo "== kernels of the default library"
SWEEP_SET=kernels timeout 500 python scripts/sweep.py run rmat1m rmat10m > $OUT/r2e_sweep_kernels.txt 2>&1; grep -E "==|us " $OUT/r2e_sweep_kernels.txt
echo "== flat variants"
SWEEP_SET=flat timeout 400 python scripts/sweep.py run rmat1m rmat10m > $OUT/r2e_sweep_flat.txt 2>&1; grep -E "==|us " $OUT/r2e_sweep_flat.txt
echo "== coo / f32"
timeout 300 python scripts/bench_formats.py coo f32 > $OUT/r2e_formats.txt 2>&1; cut -c1-220 $OUT/r2e_formats.txt
echo "== bench N=1"
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/r2e_bench.json 2> $OUT/r2e_bench.err; echo "rc=$?"; tail -n 5 $OUT/r2e_bench.err; cat $OUT/r2e_bench.json
echo "== ncu flat kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"csr_flat_kernel" -s 2 -c 1 -o $OUT/prof_r2e_flat python scripts/prof_spmv.py --impl b200 --workload rmat1m > $OUT/r2e_ncu.log 2>&1; tail -n 2 $OUT/r2e_ncu.log
