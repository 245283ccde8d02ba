set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_bricks2.txt
timeout 900 python -m pytest tests/test_gpu_bricks.py -m gpu -q -x > gpurun_out/gpu_bricks2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_bricks2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
echo done
