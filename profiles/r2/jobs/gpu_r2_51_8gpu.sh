set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_51_bench_n8.json 2> gpurun_out/r2_51_bench_n8.err; echo "rc=$?" >> gpurun_out/r2_51_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_51_bench_n4.json 2> gpurun_out/r2_51_bench_n4.err; echo "rc=$?" >> gpurun_out/r2_51_bench_n4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_51_bench_n1.json 2> gpurun_out/r2_51_bench_n1.err
echo done
