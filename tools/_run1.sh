set -x
python -m pytest tests -x -q -m gpu -k "env" > gpurun_out/pytest_env.log 2>&1; echo rc=$? >> gpurun_out/pytest_env.log
python bench.py --workload config3_env_rollouts --steps 20 --warmup 3 > gpurun_out/bench_env_pieces.json 2> gpurun_out/bench_env_pieces.err
python bench.py --workload config3_env_rollouts --steps 20 --warmup 3 --env-sequential > gpurun_out/bench_env_seq.json 2>> gpurun_out/bench_env_pieces.err
python bench.py --workload config3_env_rollouts --steps 20 --warmup 3 --math theta_carry > gpurun_out/bench_env_pieces_theta.json 2>> gpurun_out/bench_env_pieces.err
tail -3 gpurun_out/pytest_env.log; cat gpurun_out/bench_env_pieces.json gpurun_out/bench_env_seq.json gpurun_out/bench_env_pieces_theta.json | cut -c1-400
