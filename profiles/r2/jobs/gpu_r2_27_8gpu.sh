set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x -k "eight or direct" > gpurun_out/r2_27_pytest_bricks_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_27_pytest_bricks_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 20 > gpurun_out/r2_27_bench_n8_direct.json 2> gpurun_out/r2_27_bench_n8_direct.err; echo "rc=$?" >> gpurun_out/r2_27_bench_n8_direct.err
echo done
