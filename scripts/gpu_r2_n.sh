#!/bin/bash
# Round-2 call N (1 GPU, the last ~4 GPU-minutes of the round): the new GPU tests (spmv_generic.cu, strided-batch SpMM) first,
# then the whole earlier GPU suite as a regression check of the shim changes, then smoke() and a short bench line.
OUT=gpurun_out; mkdir -p $OUT
timeout 110 python -m pytest tests/test_generic_gpu.py tests/test_spmm_batched_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout=60 > $OUT/r2n_new.log 2>&1; echo "rc=$?" >> $OUT/r2n_new.log
tail -40 $OUT/r2n_new.log
timeout 140 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -q --tb=short -p no:cacheprovider --timeout=60 > $OUT/r2n_old.log 2>&1; echo "rc=$?" >> $OUT/r2n_old.log
tail -15 $OUT/r2n_old.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2n_smoke.log 2>&1; echo "rc=$?" >> $OUT/r2n_smoke.log; tail -3 $OUT/r2n_smoke.log
timeout 100 python bench.py --steps 100 --warmup 5 --no-extra > $OUT/r2n_bench.json 2> $OUT/r2n_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/r2n_bench.json
