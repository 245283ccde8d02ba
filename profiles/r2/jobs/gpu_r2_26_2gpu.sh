set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x > gpurun_out/r2_26_pytest_bricks_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_26_pytest_bricks_2gpu.log
run() { n=$1; tag=$2; shift; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 100 --warmup 20 "$@" > gpurun_out/r2_26_bench_n${n}_$tag.json 2> gpurun_out/r2_26_bench_n${n}_$tag.err; echo "rc=$?" >> gpurun_out/r2_26_bench_n${n}_$tag.err; }
run 2 peer_rebuild
AZTOT_NO_PEER_REBUILD=1 run 2 nccl_rebuild --no-halo-check
echo done
