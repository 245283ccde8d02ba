set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x -k "eight or direct or peer" > gpurun_out/r2_33_pytest_bricks_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_33_pytest_bricks_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 20 > gpurun_out/r2_33_bench_n8.json 2> gpurun_out/r2_33_bench_n8.err; echo "rc=$?" >> gpurun_out/r2_33_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 100 --warmup 20 --no-halo-check > gpurun_out/r2_33_bench_n4.json 2> gpurun_out/r2_33_bench_n4.err; echo "rc=$?" >> gpurun_out/r2_33_bench_n4.err
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_33_bench_n1.json 2> gpurun_out/r2_33_bench_n1.err
echo done
