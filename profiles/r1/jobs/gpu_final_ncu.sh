set -x
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "rc=$?" >> gpurun_out/bench18.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/prof_plain18.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches18.csv python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/ncu_launches18.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list|az_k_sort_scatter|az_k_finalize" -s 0 -c 5 -o gpurun_out/prof_step18 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/ncu_full18.log 2>&1
echo done
