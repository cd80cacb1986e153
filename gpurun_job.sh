mkdir -p gpurun_out
python tools/diag_socp2.py 2>&1 | tail -4
