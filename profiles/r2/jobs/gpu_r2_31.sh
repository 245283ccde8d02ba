set -x
mkdir -p gpurun_out
for d in 2 3; do
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 --cell-div $d > gpurun_out/r2_31_bench_celldiv$d.json 2> gpurun_out/r2_31_bench_celldiv$d.err; echo "rc=$?" >> gpurun_out/r2_31_bench_celldiv$d.err
done
echo done
