mkdir -p gpurun_out
timeout 300 python profiles/debug_fused.py > gpurun_out/debug_fused.log 2>&1
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err
