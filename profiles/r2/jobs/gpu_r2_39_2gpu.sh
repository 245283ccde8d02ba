set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x > gpurun_out/r2_39_pytest_bricks_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_39_pytest_bricks_2gpu.log
echo done
