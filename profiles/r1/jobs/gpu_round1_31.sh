mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v7.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_v7.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench35_reference.json 2> gpurun_out/bench35_reference.err; echo "rc=$?" >> gpurun_out/bench35_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_lj4m_v7.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list|az_k_finalize|az_k_sort" -s 20 -c 8 -o gpurun_out/prof_lj4m_v7 python profiles/prof_step.py lj4m 12 10 0.25 > gpurun_out/ncu_f.log 2>&1
