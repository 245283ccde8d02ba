set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 "$@" > gpurun_out/r2_11_bench_n2_$tag.json 2> gpurun_out/r2_11_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_11_bench_n2_$tag.err; }
run fused
AZTOT_NO_FUSED_HALO=1 run peer_nooverlap --no-halo-overlap --no-halo-check
run fused_b --no-halo-check
timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_11_bench_n1.json 2> gpurun_out/r2_11_bench_n1.err
echo done
