mkdir -p gpurun_out
timeout 600 python tools/sweep_config5.py --pinned > gpurun_out/r2_config5_sweep_pinned.jsonl 2> gpurun_out/r2_config5_sweep_pinned.err
cat gpurun_out/r2_config5_sweep_pinned.jsonl | cut -c95-400; tail -2 gpurun_out/r2_config5_sweep_pinned.err

