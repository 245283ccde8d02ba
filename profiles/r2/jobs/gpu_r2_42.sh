set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_42_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_42_pytest_gpu_all.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_42_bench_default_k20.json 2> gpurun_out/r2_42_bench_default_k20.err ) 2> gpurun_out/r2_42_time.txt; echo "rc=$?" >> gpurun_out/r2_42_bench_default_k20.err
for w in cs1 cs2 lj32k; do
timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_42_bench_${w}_graph.json 2> gpurun_out/r2_42_bench_${w}_graph.err
done
timeout 300 python bench.py --workload cs1 --rebuild 10 --skin 0.4 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_42_bench_cs1_r10_graph.json 2> gpurun_out/r2_42_bench_cs1_r10_graph.err
echo done
