for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err
done
python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload config3_env_rollouts --steps 20 --warmup 3 > gpurun_out/bench_scale_env_n8.json 2> gpurun_out/bench_scale_env_n8.err
for f in gpurun_out/bench_scale_*.json; do echo $f; cut -c1-260 $f; done
tail -3 gpurun_out/bench_scale_n8.err
