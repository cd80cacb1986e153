mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matpar.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/matpar_tests.log
timeout 300 python tools/time_matpar.py mpc_ltv_12_4_10 20000 2>&1 | tail -3 | tee gpurun_out/matpar_time.json
timeout 300 python tools/time_matpar.py mpc_ltv_6_3_10 20000 2>&1 | tail -3 | tee -a gpurun_out/matpar_time.json
