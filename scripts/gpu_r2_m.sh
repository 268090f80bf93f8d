#!/bin/bash
# Round-2 call M (2 GPUs): pieces of the x exchange timed alone, then bench.py with the row weight of the N = 2 fit and two
# exchange variants (SM pull kernel instead of copy-engine copies; column panels already from 2 GPUs on)
OUT=gpurun_out; mkdir -p $OUT
echo "== exchange probe"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/exchange_probe.py > $OUT/r2m_probe.json 2> $OUT/r2m_probe.err; echo "rc=$?"; cat $OUT/r2m_probe.json | cut -c1-900
bash scripts/gpu_r2_multi.sh 2 "w8:" "sm64:--no-extra:B200SPMV_XCHG_SM_CTAS=64" "sm148:--no-extra:B200SPMV_XCHG_SM_CTAS=148" "panels2:--no-extra:B200SPMV_PANELS_FROM=2"
