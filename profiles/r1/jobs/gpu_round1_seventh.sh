set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_long.py -m gpu -q -x -s > gpurun_out/gpu_long1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_long1.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests8.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "rc=$?" >> gpurun_out/bench9.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err; echo "rc=$?" >> gpurun_out/bench9_ref.err
echo done
