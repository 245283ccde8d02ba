set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_44_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_44_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_44_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_44_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_44_bench_reference.json 2> gpurun_out/r2_44_bench_reference.err; echo "rc=$?" >> gpurun_out/r2_44_bench_reference.err
timeout 900 python bench.py > gpurun_out/r2_44_bench_default.json 2> gpurun_out/r2_44_bench_default.err; echo "rc=$?" >> gpurun_out/r2_44_bench_default.err
echo done
