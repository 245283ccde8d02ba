set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 20 --no-halo-check "$@" > gpurun_out/r2_38_bench_n8_$tag.json 2> gpurun_out/r2_38_bench_n8_$tag.err; echo "rc=$?" >> gpurun_out/r2_38_bench_n8_$tag.err; }
run direct
AZTOT_NO_FUSED_HALO=1 run peer_nooverlap --no-halo-overlap
AZTOT_NO_FUSED_HALO=1 run peer_overlap
echo done
