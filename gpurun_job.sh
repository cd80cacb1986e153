mkdir -p gpurun_out
python tools/sweep_batch.py 2>&1 | tail -4 | tee gpurun_out/sweep_batch.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mpc_12 or full_size" 2>&1 | tail -3
