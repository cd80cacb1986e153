mkdir -p gpurun_out
python gpurun_diag.py 2>&1 | grep -v Warning | tail -30
python -m pytest tests -m gpu -q 2>&1 | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:admm_ -s 6 -c 2 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 50000 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
