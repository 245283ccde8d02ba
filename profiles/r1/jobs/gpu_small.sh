mkdir -p gpurun_out
for w in cs1 cs2 lj32k; do
timeout 200 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "rc=$?" >> gpurun_out/bench_$w.err
done
timeout 200 python bench.py --workload cs1 --rebuild 10 --skin 1.0 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_cs1_r10.json 2> gpurun_out/bench_cs1_r10.err
timeout 200 python bench.py --workload lj32k --rebuild 10 --skin 0.4 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_lj32k_r10.json 2> gpurun_out/bench_lj32k_r10.err
timeout 600 python -m pytest tests/test_gpu_long.py -m gpu -q -x -s -k "driver" > gpurun_out/gpu_long4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_long4.log
