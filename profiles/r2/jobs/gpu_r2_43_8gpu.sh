set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x > gpurun_out/r2_43_pytest_bricks_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_43_pytest_bricks_8gpu.log
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_43_bench_n$n.json 2> gpurun_out/r2_43_bench_n$n.err; echo "rc=$?" >> gpurun_out/r2_43_bench_n$n.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_43_bench_n1.json 2> gpurun_out/r2_43_bench_n1.err
echo done
