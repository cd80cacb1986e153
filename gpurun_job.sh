mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/sweep_variants.py run 2>&1 | tail -5
python tools/sweep_batch.py 2>&1 | tail -5 | tee gpurun_out/sweep_batch.jsonl
