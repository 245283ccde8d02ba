set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 --no-halo-check "$@" > gpurun_out/r2_12_bench_n2_$tag.json 2> gpurun_out/r2_12_bench_n2_$tag.err; echo "rc=$?" >> gpurun_out/r2_12_bench_n2_$tag.err; }
AZTOT_PUT_DBG=1 run dbg1
AZTOT_PUT_DBG=2 run dbg2
AZTOT_PUT_DBG=3 run dbg3
run fused
for v in wm6 wm8; do for w in cs1 cs2 lj32k; do
AZTOT_B200_LIB=$PWD/aztotmd_b200/lib/variants/$v/libaztot_b200.so timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_12_bench_${w}_$v.json 2> gpurun_out/r2_12_bench_${w}_$v.err
done; done
for w in cs1 cs2 lj32k; do timeout 300 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 2 --graph > gpurun_out/r2_12_bench_${w}_wm1.json 2> gpurun_out/r2_12_bench_${w}_wm1.err; done
echo done
