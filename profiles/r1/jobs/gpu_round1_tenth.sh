set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long.py -m gpu -q -x > gpurun_out/gpu_tests11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests11.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "rc=$?" >> gpurun_out/bench13.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench13b.json 2> gpurun_out/bench13b.err; echo "rc=$?" >> gpurun_out/bench13b.err
timeout 200 python bench.py --rebuild 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench13_r1.json 2> gpurun_out/bench13_r1.err
timeout 200 python bench.py --workload cs1 --rebuild 10 --skin 1.0 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench13_cs1_r10.json 2> gpurun_out/bench13_cs1_r10.err
echo done
