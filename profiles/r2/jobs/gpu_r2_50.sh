set -x
mkdir -p gpurun_out
# split-mirror gathers of the plain LJ list pass (LDG.64 + LDG.32 instead of LDG.128): on (default) against off (--option soa_gather=0), same binary, same job
B="--steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2"
timeout 300 python bench.py $B > gpurun_out/r2_50_bench_soa_on_a.json 2> gpurun_out/r2_50_bench_soa_on_a.err; echo "rc=$?" >> gpurun_out/r2_50_bench_soa_on_a.err
timeout 300 python bench.py $B --option soa_gather=0 > gpurun_out/r2_50_bench_soa_off_a.json 2> gpurun_out/r2_50_bench_soa_off_a.err; echo "rc=$?" >> gpurun_out/r2_50_bench_soa_off_a.err
timeout 300 python bench.py $B > gpurun_out/r2_50_bench_soa_on_b.json 2> gpurun_out/r2_50_bench_soa_on_b.err; echo "rc=$?" >> gpurun_out/r2_50_bench_soa_on_b.err
timeout 300 python bench.py $B --option soa_gather=0 > gpurun_out/r2_50_bench_soa_off_b.json 2> gpurun_out/r2_50_bench_soa_off_b.err; echo "rc=$?" >> gpurun_out/r2_50_bench_soa_off_b.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_50_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_50_pytest_gpu_all.log
echo done
