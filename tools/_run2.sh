timeout 300 python -m pytest tests -x -q -m gpu -k "actor or ppo or rollout_policy or bptt" > gpurun_out/pytest_actor.log 2>&1; echo rc=$? >> gpurun_out/pytest_actor.log
tail -5 gpurun_out/pytest_actor.log
timeout 120 python bench.py --workload config3_actor_rollouts --steps 5 --warmup 3 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("actor", d["ms_per_step"], d["value"]/1e9)'
