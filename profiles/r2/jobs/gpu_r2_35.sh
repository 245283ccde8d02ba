set -x
mkdir -p gpurun_out
timeout 300 python profiles/prof_step.py lj4m 22 10 0.25 0 > gpurun_out/r2_35_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list|az_k_sort_scatter|az_k_finalize" -s 8 -c 6 -o gpurun_out/r2_35_full python profiles/prof_step.py lj4m 22 10 0.25 0 > gpurun_out/r2_35_ncu_full.log 2>&1
echo done
