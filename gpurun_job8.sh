mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python tools/sweep_config5.py > gpurun_out/r2_config5_sweep.jsonl 2> gpurun_out/r2_config5_sweep.err
cat gpurun_out/r2_config5_sweep.jsonl | cut -c100-400; tail -2 gpurun_out/r2_config5_sweep.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -2 gpurun_out/r2_bench_n8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('mpc N=8', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('parity') if 'mat' in k else '')
PY
