set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_23_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_23_pytest_parity.log
for w in cs1 cs2 lj32k; do
timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_23_bench_${w}_graph.json 2> gpurun_out/r2_23_bench_${w}_graph.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2_23_launches_cs2.csv python bench.py --workload cs2 --rebuild 1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_23_ncu_cs2.log 2>&1
echo done
