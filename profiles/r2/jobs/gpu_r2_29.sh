set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_29_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_29_pytest_gpu_all.log
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_29_bench_ovl.json 2> gpurun_out/r2_29_bench_ovl.err; echo "rc=$?" >> gpurun_out/r2_29_bench_ovl.err
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 --option finalize_overlap=0 > gpurun_out/r2_29_bench_noovl.json 2> gpurun_out/r2_29_bench_noovl.err; echo "rc=$?" >> gpurun_out/r2_29_bench_noovl.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_29_bench_ovl_k20.json 2> gpurun_out/r2_29_bench_ovl_k20.err
echo done
