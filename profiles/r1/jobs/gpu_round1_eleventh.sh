set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_bricks.py -m gpu -q -x > gpurun_out/gpu_bricks5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_bricks5.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_v4.json 2> gpurun_out/bench_n2_v4.err; echo "rc=$?" >> gpurun_out/bench_n2_v4.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --gpus 1 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "rc=$?" >> gpurun_out/bench14.err
