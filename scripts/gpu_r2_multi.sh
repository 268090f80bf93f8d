#!/bin/bash
# Round-2 multi-GPU call: bench.py at N GPUs; compares exchange mechanisms / row-partition work models.
# usage: gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N "tag:args" "tag:args:VAR=VALUE" ...'
N=${1:-2}; shift
OUT=gpurun_out; mkdir -p $OUT
run() {  # tag, extra args, optional VAR=VALUE for the environment
  env $3 BENCH_WATCHDOG_S=170 timeout 220 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus $N --steps 50 --warmup 5 $2 > $OUT/r2_n${N}_$1.json 2> $OUT/r2_n${N}_$1.err
  echo "== $1 rc=$?"; grep -v "^\*\|OMP_NUM\|NCCL version\|^$" $OUT/r2_n${N}_$1.err | tail -n 3 | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r2_n${N}_$1.json"))
    pr = d["roofline"]["per_rank"]
    print("value", d["value"], "GB/s  ms/step", d["ms_per_step"], " e2e", d["e2e"]["ms_per_step"], " local/rank", pr.get("local_product_us_per_rank"), " rows/rank", pr.get("rows_per_rank"), " xchg", pr.get("exchange_alone_us"))
    print("   ", d["exchange"][:100], "| cg", (d.get("cg_config4") or {}).get("value"))
except Exception as e:
    print("no json:", e)
PY
}
for spec in "$@"; do   # "tag:bench args" or "tag:bench args:VAR=VALUE"
  tag="${spec%%:*}"; rest="${spec#*:}"
  if [[ "$rest" == *:* ]]; then run "$tag" "${rest%%:*}" "${rest#*:}"; else run "$tag" "$rest" "B200_NOOP=1"; fi
done
