set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_14_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_14_pytest_parity.log
timeout 600 python bench.py --steps 100 --warmup 20 --no-small-deck > gpurun_out/r2_14_bench_tlist.json 2> gpurun_out/r2_14_bench_tlist.err; echo "rc=$?" >> gpurun_out/r2_14_bench_tlist.err
timeout 600 python bench.py --steps 100 --warmup 20 --option tile_list=0 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_14_bench_rows32.json 2> gpurun_out/r2_14_bench_rows32.err; echo "rc=$?" >> gpurun_out/r2_14_bench_rows32.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_14_prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_14_launches.csv python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_14_ncu_launches.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_tlist|az_k_build_tlist" -s 0 -c 3 -o gpurun_out/r2_14_full python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_14_ncu_full.log 2>&1
echo done
