set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_bricks.py -m gpu -q -x > gpurun_out/gpu_bricks6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_bricks6.log
for n in 8 4 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/bench_n${n}_v5.json 2> gpurun_out/bench_n${n}_v5.err; echo "rc=$?" >> gpurun_out/bench_n${n}_v5.err
done
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1_v5.json 2> gpurun_out/bench_n1_v5.err
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/bench_n8_ref.json 2> gpurun_out/bench_n8_ref.err
