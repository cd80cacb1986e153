set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
tail -c 600 gpurun_out/r2_bench_n2_final.json; tail -3 gpurun_out/r2_bench_n2_final.err
timeout 300 python -m pytest tests/test_multi_device_api.py -x -q -m gpu 2>&1 | tail -3
