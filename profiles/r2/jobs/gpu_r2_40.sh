set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_40_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_40_pytest_gpu_all.log
for w in cs1 cs2 lj32k; do
timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_40_bench_${w}_graph.json 2> gpurun_out/r2_40_bench_${w}_graph.err
timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph --option pdl=0 > gpurun_out/r2_40_bench_${w}_graph_nopdl.json 2> gpurun_out/r2_40_bench_${w}_graph_nopdl.err
done
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_40_bench_cs1_plain.json 2> gpurun_out/r2_40_bench_cs1_plain.err
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --option pdl=0 > gpurun_out/r2_40_bench_cs1_plain_nopdl.json 2> gpurun_out/r2_40_bench_cs1_plain_nopdl.err
echo done
