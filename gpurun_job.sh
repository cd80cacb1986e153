mkdir -p gpurun_out
python tools/diag_socp.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_socp_ipm.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/socp_tests.log
timeout 300 python tools/bench_socp.py --batch 20000 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/socp_bench.log
