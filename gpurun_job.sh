python tools/sweep_variants.py run 2>&1 | tail -4
