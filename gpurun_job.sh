mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_socp_ipm.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/socp_tests.log
