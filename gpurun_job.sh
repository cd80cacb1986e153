mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_matpar.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests_matpar.log
timeout 300 python tools/time_matpar.py mpc_ltv_6_3_10 20000 2>&1 | tail -1
