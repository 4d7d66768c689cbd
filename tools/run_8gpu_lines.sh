#!/bin/bash
# N-GPU bench lines of the sharded workloads (torchrun, NCCL), appended to gpurun_out/bench_${N}gpu.jsonl
N=${1:-8}
OUT=gpurun_out/bench_${N}gpu.jsonl
: > $OUT
port=29520
WLS=${2:-"config2_batched_icem config3_env_rollouts config3_actor_rollouts config3_collect_experience"}
for wl in $WLS; do
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --workload $wl --steps 20 --warmup 3 --no-cpu-baseline >> $OUT 2>> gpurun_out/bench_${N}gpu.err
done
python - <<PY
import json
for line in open("$OUT"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["config"]["workload"], d["n_gpus"], "%.4g" % d["value"], d["unit"], "ms %.4g" % d["ms_per_step"], "e2e %.4g" % (d.get("e2e") or {}).get("value", 0))
PY
