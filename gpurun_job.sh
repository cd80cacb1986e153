mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v4.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
ncu --set full --clock-control none --import-source on -k regex:admm_ -s 6 -c 2 -o gpurun_out/prof_v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 50000 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
