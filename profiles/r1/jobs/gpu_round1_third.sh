set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests3.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench3_list.json 2> gpurun_out/bench3.err; echo "bench rc=$?" >> gpurun_out/bench3.err
timeout 300 python bench.py --steps 100 --warmup 10 --pair-variant 1 --no-cpu-baseline > gpurun_out/bench3_tiles.json 2>> gpurun_out/bench3.err
timeout 300 python bench.py --steps 100 --warmup 10 --rebuild 1 --no-cpu-baseline > gpurun_out/bench3_rebuild1.json 2>> gpurun_out/bench3.err
timeout 300 python bench.py --steps 200 --warmup 20 --skin 0.25 --no-cpu-baseline > gpurun_out/bench3_list_skin025.json 2>> gpurun_out/bench3.err
timeout 300 python profiles/prof_step.py lj4m 12 > gpurun_out/prof_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches3.csv python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_launches3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:az_k_pair -s 0 -c 3 -o gpurun_out/prof_pair3 python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_full3.log 2>&1
echo done
