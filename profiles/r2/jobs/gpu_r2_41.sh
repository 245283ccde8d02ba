set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_41_bench_default_k20.json 2> gpurun_out/r2_41_bench_default_k20.err ) 2> gpurun_out/r2_41_time.txt; echo "rc=$?" >> gpurun_out/r2_41_bench_default_k20.err
echo done
