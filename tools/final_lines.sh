#!/bin/bash
# Final single-GPU evidence of a round: GPU tests, smoke, the default bench + its CPU arm, one line per other workload.
R=${1:-r01e}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${R}.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_${R}.log
tail -3 gpurun_out/pytest_gpu_${R}.log
python __graft_entry__.py smoke > gpurun_out/smoke_${R}.log 2>&1; echo rc=$? >> gpurun_out/smoke_${R}.log; tail -4 gpurun_out/smoke_${R}.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_config2_reference.json 2> gpurun_out/bench_${R}.err
python bench.py > gpurun_out/bench_${R}_config2.json 2>> gpurun_out/bench_${R}.err
: > gpurun_out/bench_${R}_other.jsonl
for wl in config1_closed_loop config2_colored config4_ensemble_icem bptt_rollout_grad config3_collect_experience; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 >> gpurun_out/bench_${R}_other.jsonl 2>> gpurun_out/bench_${R}.err
done
python - <<PY
import json
for f in ("gpurun_out/bench_${R}_config2_reference.json", "gpurun_out/bench_${R}_config2.json", "gpurun_out/bench_${R}_other.jsonl"):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print(d.get("impl", "ours"), d["config"].get("workload"), "%.4g" % d["value"], d["unit"], "ms %.4g" % d["ms_per_step"],
                  "e2e %.4g" % ((d.get("e2e") or {}).get("value") or 0), "frac %s" % ((d.get("roofline") or {}).get("frac")))
PY
tail -3 gpurun_out/bench_${R}.err
