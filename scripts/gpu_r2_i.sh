#!/bin/bash
# Round-2 call I: flat kernel (gap fill in the chunk post-pass; look-back on/off), csr_short_kernel, SpMM vs closed library
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (csr kernels, flat, short, sharded-free subset)"
timeout 600 python -m pytest tests -m gpu -x -q -k "csr_kernel or flat or short or plan or edge or stream or graph or smoke or solver_samples or reference_sample" > $OUT/r2i_pytest.log 2>&1; echo "rc=$?"; tail -n 5 $OUT/r2i_pytest.log
echo "== flat2 variants"
SWEEP_SET=flat2 timeout 500 python scripts/sweep.py run rmat1m rmat10m rmat1m_f32 > $OUT/r2i_sweep_flat2.txt 2>&1; grep -E "==|us " $OUT/r2i_sweep_flat2.txt
echo "== short variants"
SWEEP_SET=short timeout 500 python scripts/sweep.py run stencil5_8192 laplace7_256_f32 uniform1m > $OUT/r2i_sweep_short.txt 2>&1; grep -E "==|us " $OUT/r2i_sweep_short.txt
echo "== kernels of the default library on the short-row matrices"
SWEEP_SET=kernels timeout 500 python scripts/sweep.py run stencil5_8192 laplace7_256_f32 uniform1m > $OUT/r2i_sweep_kernels_short.txt 2>&1; grep -E "==|us " $OUT/r2i_sweep_kernels_short.txt
echo "== formats (incl. SpMM config 5)"
timeout 500 python scripts/bench_formats.py > $OUT/r2i_formats.txt 2>&1; cut -c1-260 $OUT/r2i_formats.txt; cp $OUT/bench_formats.json $OUT/r2i_formats.json
