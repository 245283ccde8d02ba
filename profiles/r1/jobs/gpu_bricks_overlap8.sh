set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_n8_ov.json 2> gpurun_out/bench_n8_ov.err; echo "rc=$?" >> gpurun_out/bench_n8_ov.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 100 --warmup 10 --no-halo-overlap > gpurun_out/bench_n8_noov.json 2> gpurun_out/bench_n8_noov.err; echo "rc=$?" >> gpurun_out/bench_n8_noov.err
