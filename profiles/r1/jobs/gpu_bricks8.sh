set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_bricks8.txt
timeout 300 python -m pytest tests/test_gpu_bricks.py -m gpu -q -x -k "eight or four or distributed" > gpurun_out/gpu_bricks8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_bricks8.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?" >> gpurun_out/bench_n8.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 50 --warmup 10 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?" >> gpurun_out/bench_n4.err
echo done
