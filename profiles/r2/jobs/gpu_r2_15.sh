set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 20 --no-small-deck --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_15_bench_tlist.json 2> gpurun_out/r2_15_bench_tlist.err; echo "rc=$?" >> gpurun_out/r2_15_bench_tlist.err
timeout 300 python profiles/prof_step.py lj4m 22 10 0.25 0 > gpurun_out/r2_15_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_tlist|az_k_build_tlist" -s 4 -c 3 -o gpurun_out/r2_15_full python profiles/prof_step.py lj4m 22 10 0.25 0 > gpurun_out/r2_15_ncu_full.log 2>&1
echo done
