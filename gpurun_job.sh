mkdir -p gpurun_out
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_mpc_n$N.json
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --workload portfolio_socp 2>>gpurun_out/bench_n8.err | tail -1 | tee gpurun_out/bench_socp_n8.json
tail -3 gpurun_out/bench_n8.err
