for k in 1 2 3; do
timeout 300 python bench.py --steps 200 --warmup 20 --skin 0.25 --no-cpu-baseline > gpurun_out/dbg_bench_$k.json 2> gpurun_out/dbg_bench_$k.err; echo "rc=$?" >> gpurun_out/dbg_bench_$k.err
done
