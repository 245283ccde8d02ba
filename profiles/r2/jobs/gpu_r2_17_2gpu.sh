set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 --no-halo-check "$@" > gpurun_out/r2_17_bench_n2_$tag.json 2> gpurun_out/r2_17_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_17_bench_n2_$tag.err; }
run direct
run nooverlap --no-halo-overlap
run direct_r20 --rebuild 20 --skin 0.5
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_17_bench_n1_gpu1.json 2> gpurun_out/r2_17_bench_n1_gpu1.err
echo done
