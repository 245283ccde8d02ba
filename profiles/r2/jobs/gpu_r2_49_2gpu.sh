set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x > gpurun_out/r2_49_pytest_bricks_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_49_pytest_bricks_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_49_bench_n2.json 2> gpurun_out/r2_49_bench_n2.err; echo "rc=$?" >> gpurun_out/r2_49_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_49_bench_ref_n2.json 2> gpurun_out/r2_49_bench_ref_n2.err; echo "rc=$?" >> gpurun_out/r2_49_bench_ref_n2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_49_bench_n1.json 2> gpurun_out/r2_49_bench_n1.err
echo done
