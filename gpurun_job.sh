mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --with-grad > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"admm_|qp_grad" -s 6 -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --with-grad > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
