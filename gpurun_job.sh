mkdir -p gpurun_out
for T in 384 512; do
timeout 300 python tools/bench_socp.py --code-dir tools/_variants/socp_T$T --batch 20000 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/socp_bench_T$T.log
done
