python bench.py --workload config3_actor_rollouts --steps 10 --warmup 3 > gpurun_out/bench_actor_tc.json 2> gpurun_out/bench_actor_tc.err; cut -c1-300 gpurun_out/bench_actor_tc.json
python bench.py --workload config3_actor_rollouts --steps 10 --warmup 3 --actor-kernel cuda_cores > gpurun_out/bench_actor_cc.json 2>> gpurun_out/bench_actor_tc.err; cut -c1-200 gpurun_out/bench_actor_cc.json
python bench.py --workload bptt_rollout_grad --steps 20 --warmup 3 > gpurun_out/bench_bptt2.json 2>> gpurun_out/bench_actor_tc.err; cut -c1-260 gpurun_out/bench_bptt2.json
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_h.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_h.log; tail -3 gpurun_out/pytest_gpu_h.log
