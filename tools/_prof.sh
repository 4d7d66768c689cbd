python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_e.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_e.log; tail -3 gpurun_out/pytest_gpu_e.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_e1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icem_plan -s 3 -c 1 -o gpurun_out/prof_plan_e -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e2.log 2>&1
tail -2 gpurun_out/ncu_e2.log | cut -c1-200
