set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_long.py -m gpu -q -x -s > gpurun_out/r2_37_pytest_long.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_37_pytest_long.log
echo done
