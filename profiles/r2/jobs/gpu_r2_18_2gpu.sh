set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 "$@" > gpurun_out/r2_18_bench_n2_$tag.json 2> gpurun_out/r2_18_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_18_bench_n2_$tag.err; }
run direct
run direct200 --steps 200 --no-halo-check
timeout 600 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x -k "peer or verlet or rider" > gpurun_out/r2_18_pytest_bricks.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_18_pytest_bricks.log
echo done
