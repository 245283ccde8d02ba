set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x -s > gpurun_out/r2_09_pytest_bricks_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_09_pytest_bricks_2gpu.log
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 "$@" > gpurun_out/r2_09_bench_n2_$tag.json 2> gpurun_out/r2_09_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_09_bench_n2_$tag.err; }
run fused
AZTOT_NO_FUSED_HALO=1 run peer_nooverlap --no-halo-overlap --no-halo-check
run fused_b --no-halo-check
timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_09_bench_n1.json 2> gpurun_out/r2_09_bench_n1.err
echo done
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_09_bench_cs1.json 2> gpurun_out/r2_09_bench_cs1.err
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --option warp_prune=0 > gpurun_out/r2_09_bench_cs1_noprune.json 2> gpurun_out/r2_09_bench_cs1_noprune.err
timeout 300 python bench.py --workload lj32k --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_09_bench_lj32k.json 2> gpurun_out/r2_09_bench_lj32k.err
timeout 300 python bench.py --workload lj32k --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --option warp_prune=0 > gpurun_out/r2_09_bench_lj32k_noprune.json 2> gpurun_out/r2_09_bench_lj32k_noprune.err
timeout 300 python bench.py --workload cs2 --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_09_bench_cs2.json 2> gpurun_out/r2_09_bench_cs2.err
