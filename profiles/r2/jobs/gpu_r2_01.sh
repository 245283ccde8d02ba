set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_01_smi.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_01_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_01_pytest_parity.log
timeout 400 python bench.py --steps 100 --warmup 20 > gpurun_out/r2_01_bench_default.json 2> gpurun_out/r2_01_bench_default.err; echo "rc=$?" >> gpurun_out/r2_01_bench_default.err
for v in r1 w64 w40; do
AZTOT_B200_LIB=$PWD/aztotmd_b200/lib/variants/$v/libaztot_b200.so timeout 400 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_01_bench_$v.json 2> gpurun_out/r2_01_bench_$v.err; echo "rc=$?" >> gpurun_out/r2_01_bench_$v.err
done
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_01_prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_01_launches.csv python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_01_ncu_launches.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list" -s 0 -c 3 -o gpurun_out/r2_01_full python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_01_ncu_full.log 2>&1
echo done
