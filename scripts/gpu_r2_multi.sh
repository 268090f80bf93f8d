#!/bin/bash
# Round-2 multi-GPU call: bench.py at N GPUs with the peer-copy exchange, then with the NCCL all-gather for comparison.
# usage: gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N'
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo_n$N.txt 2>&1
run() {  # tag, extra args
  BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus $N --steps 50 --warmup 5 $2 > $OUT/r2_n${N}_$1.json 2> $OUT/r2_n${N}_$1.err
  echo "== $1 rc=$?"; tail -n 3 $OUT/r2_n${N}_$1.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r2_n${N}_$1.json"))
    print("value", d["value"], "GB/s  ms/step", d["ms_per_step"], " kernel_us", d["roofline"]["kernel_avg_us"], " e2e", d["e2e"]["ms_per_step"], "|", d["exchange"][:160])
    print("cg", d.get("cg_config4"))
except Exception as e:
    print("no json:", e)
PY
}
run p2p "--exchange auto"
run allgather "--exchange allgather --no-extra"
if [ "$N" -le 2 ]; then
B200SPMV_STEP_GRAPH=0 run p2p_nograph "--exchange auto --no-extra"
fi
