set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_long.py -m gpu -q -x -s > gpurun_out/gpu_long2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_long2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests9.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "rc=$?" >> gpurun_out/bench10.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/prof_plain10.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches10.csv python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/ncu_launches10.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list|az_k_integrate_ba" -s 0 -c 3 -o gpurun_out/prof_step10 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/ncu_full10.log 2>&1
echo done
