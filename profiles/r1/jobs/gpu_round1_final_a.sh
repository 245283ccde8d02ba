set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q -x > gpurun_out/gpu_all1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_all1.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke1.log
timeout 400 python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "rc=$?" >> gpurun_out/bench12.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err
timeout 300 python bench.py --workload nacl1m --rebuild 5 --skin 0.6 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench12_nacl1m.json 2> gpurun_out/bench12_nacl1m.err; echo "rc=$?" >> gpurun_out/bench12_nacl1m.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/prof_plain12.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches12.csv python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/ncu_launches12.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list|az_k_integrate_ba|az_k_sort_scatter" -s 0 -c 4 -o gpurun_out/prof_step12 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/ncu_full12.log 2>&1
echo done
