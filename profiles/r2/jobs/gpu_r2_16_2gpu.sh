set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_16_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x > gpurun_out/r2_16_pytest_bricks.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_16_pytest_bricks.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 "$@" > gpurun_out/r2_16_bench_n2_$tag.json 2> gpurun_out/r2_16_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_16_bench_n2_$tag.err; }
run direct
AZTOT_NO_DIRECT_HALO=1 run fused --no-halo-check
run nooverlap --no-halo-overlap --no-halo-check
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_16_bench_n1.json 2> gpurun_out/r2_16_bench_n1.err
echo done
