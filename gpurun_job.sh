mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log
timeout 600 python bench.py --workload mpc_ltv --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ltv.json
timeout 600 python bench.py --steps 10 --warmup 3 --with-grad 2>&1 | tail -1 > gpurun_out/bench_mpc.json
python -c "
import json
for f in ('bench_ltv','bench_mpc'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['config'].get('gradient'), d['cpu_baseline']['value'])
"
