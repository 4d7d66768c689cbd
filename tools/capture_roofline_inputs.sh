#!/bin/bash
# Re-capture the ncu inputs of the headline roofline (profiles/roofline_inputs.json) on the GPU box.
#   usage (here):  gpurun --timeout 900 -- 'bash tools/capture_roofline_inputs.sh r02o'
# 1. the plain command must exit 0 first; 2. one launch of the fused plan kernel under `ncu --set full` (a number printed
# under ncu is never a bench value); 3. the metric subset (tools/ncu_summary.py) and the launch list of the default
# bench command go to gpurun_out/ -- copy them into profiles/ and update roofline_inputs.json from the summary:
#   executed_lane_instr_per_transition = smsp__inst_executed.sum x 32 / (4096 x 527 x 30 x 5)
#   dram_bytes_per_launch              = dram__bytes_read.sum + dram__bytes_write.sum
R=${1:-r02o}
set -e
python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/${R}_plain.json
ncu --set full --clock-control none --import-source on -k regex:icem_plan_pendulum_kernel -s 3 -c 1 -f \
    -o gpurun_out/${R}_plan_kernel python bench.py --steps 1 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/${R}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_plan_kernel.ncu-rep > gpurun_out/${R}_plan_kernel_ncu_full.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
grep -E "smsp__inst_executed.sum|dram__bytes_(read|write).sum|gpu__time_duration.sum|smsp__issue_active" gpurun_out/${R}_plan_kernel_ncu_full.txt
