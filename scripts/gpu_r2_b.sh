#!/bin/bash
# Round-2 call B: seg kernels with run-time step loops (instruction-cache fix)
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (seg / coo)"
timeout 600 python -m pytest tests -m gpu -x -q -k "every_csr_kernel or coo" > $OUT/r2b_pytest.log 2>&1; echo "rc=$?"; tail -n 5 $OUT/r2b_pytest.log
echo "== kernels of the default library"
SWEEP_SET=kernels timeout 500 python scripts/sweep.py run rmat1m rmat10m uniform1m stencil5_4096 > $OUT/r2b_sweep_kernels.txt 2>&1; grep -E "==|us " $OUT/r2b_sweep_kernels.txt
echo "== seg variants on rmat1m"
SWEEP_SET=seg timeout 400 python scripts/sweep.py run rmat1m > $OUT/r2b_sweep_seg.txt 2>&1; grep -E "==|us " $OUT/r2b_sweep_seg.txt
echo "== coo / f32"
timeout 300 python scripts/bench_formats.py coo f32 > $OUT/r2b_formats.txt 2>&1; cut -c1-220 $OUT/r2b_formats.txt
echo "== ncu seg kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"csr_seg_kernel" -s 2 -c 1 -o $OUT/prof_r2b_seg python scripts/prof_spmv.py --impl b200 --workload rmat1m > $OUT/r2b_ncu.log 2>&1; tail -n 2 $OUT/r2b_ncu.log
