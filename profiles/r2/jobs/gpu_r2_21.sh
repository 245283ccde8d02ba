set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload cs1 --rebuild 1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_21_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_warp" -s 3 -c 1 -o gpurun_out/r2_21_full_cs1 python bench.py --workload cs1 --rebuild 1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_21_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_warp" -s 3 -c 1 -o gpurun_out/r2_21_full_cs2 python bench.py --workload cs2 --rebuild 1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_21_ncu_full2.log 2>&1
echo done
