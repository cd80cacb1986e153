python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/sweep_variants.py run 2>&1 | tail -4
