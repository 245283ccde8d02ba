set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_48_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_48_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_48_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_48_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_48_bench_reference.json 2> gpurun_out/r2_48_bench_reference.err; echo "rc=$?" >> gpurun_out/r2_48_bench_reference.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_48_bench_k20.json 2> gpurun_out/r2_48_bench_k20.err ) 2> gpurun_out/r2_48_time.txt; echo "rc=$?" >> gpurun_out/r2_48_bench_k20.err
echo done
