#!/bin/bash
# The multi-GPU lines of one box: usage tools/multi_gpu_lines.sh N  (under gpurun --gpus N).  One JSON line per run
# into gpurun_out/r02_${N}gpu.jsonl: config 2 (weak), config 5 sweep (strong per point, CPU column), config 3 env
# rollouts and policy-in-the-loop rollouts in both scaling modes.
N=$1
OUT=gpurun_out/r02_${N}gpu.jsonl
: > $OUT
run() {
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" >> $OUT 2>> gpurun_out/r02_${N}gpu.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
         bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/r02_${N}gpu.err; fi
  echo "rc=$? $*"
}
run --steps 20 --warmup 3 --no-others
run --workload config5_sweep
run --workload config3_env_rollouts --scaling strong --steps 10
run --workload config3_env_rollouts --scaling weak --steps 10
run --workload config3_actor_rollouts --scaling strong --steps 5
run --workload config3_actor_rollouts --scaling weak --steps 5
run --workload config3_collect_experience --steps 5
wc -l $OUT
