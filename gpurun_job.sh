mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_socp_ipm.py -m gpu -x -q -k "network" 2>&1 | tail -5 | tee gpurun_out/gpu_tests_network.log
