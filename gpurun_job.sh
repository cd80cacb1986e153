mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --with-grad 2>&1 | tail -1 > gpurun_out/bench_v6.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('gradient'))"
