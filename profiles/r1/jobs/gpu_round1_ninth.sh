set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_long.py -m gpu -q -x -s > gpurun_out/gpu_long3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_long3.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "rc=$?" >> gpurun_out/bench11.err
echo done
