mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_socp_ipm.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/socp_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm -c 1 -o gpurun_out/ipm_r1_v5 -f python tools/bench_socp.py --batch 1184 --steps 1 --warmup 0 --cpu-sample 0 2>&1 | tail -1 | cut -c1-100
