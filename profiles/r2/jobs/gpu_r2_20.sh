set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_20_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_20_pytest_parity.log
for w in cs1 cs2 lj32k; do
timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_20_bench_${w}_graph.json 2> gpurun_out/r2_20_bench_${w}_graph.err
done
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph --pair-variant 3 > gpurun_out/r2_20_bench_cs1_graph_warp.json 2> gpurun_out/r2_20_bench_cs1_graph_warp.err
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_20_bench_cs1_plain.json 2> gpurun_out/r2_20_bench_cs1_plain.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_20_launches_cs1.csv python bench.py --workload cs1 --rebuild 1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_20_ncu_cs1.log 2>&1
echo done
