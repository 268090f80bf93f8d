#!/bin/bash
# Round-2 call A: new seg kernels under the parity fixture, kernel sweep, micro-benchmarks.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
echo "== pytest (new code paths)"
timeout 600 python -m pytest tests -m gpu -x -q -k "every_csr_kernel or coo or plan_is_trusted or toy or edge or solver or sample or forwarded" > $OUT/r2a_pytest.log 2>&1; echo "rc=$?"; tail -n 15 $OUT/r2a_pytest.log
echo "== kernels of the default library"
SWEEP_SET=kernels timeout 500 python scripts/sweep.py run rmat1m uniform1m stencil5_4096 > $OUT/r2a_sweep_kernels.txt 2>&1; grep -E "==|us " $OUT/r2a_sweep_kernels.txt
echo "== seg variants on rmat1m"
SWEEP_SET=seg timeout 400 python scripts/sweep.py run rmat1m > $OUT/r2a_sweep_seg.txt 2>&1; grep -E "==|us " $OUT/r2a_sweep_seg.txt
echo "== coo / f32"
timeout 300 python scripts/bench_formats.py coo f32 > $OUT/r2a_formats.txt 2>&1; cut -c1-220 $OUT/r2a_formats.txt
echo "== micro_gather on the real index stream"
python -c "
from cudalibrarysamples_b200 import workloads as W
off,col,val=W.rmat_csr(1000000); col.cpu().numpy().tofile('/tmp/rmat_col.bin')" && timeout 200 scripts/micro_gather /tmp/rmat_col.bin 1000000 > $OUT/r2a_micro_hot.txt 2>&1; cat $OUT/r2a_micro_hot.txt
echo "== ncu seg kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"csr_seg_kernel" -s 2 -c 1 -o $OUT/prof_r2a_seg python scripts/prof_spmv.py --impl b200 --workload rmat1m > $OUT/r2a_ncu.log 2>&1; tail -n 2 $OUT/r2a_ncu.log
