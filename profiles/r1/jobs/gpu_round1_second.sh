set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests2.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench2_list.json 2> gpurun_out/bench2.err; echo "bench rc=$?" >> gpurun_out/bench2.err
timeout 300 python bench.py --steps 100 --warmup 10 --pair-variant 1 --no-cpu-baseline > gpurun_out/bench2_tiles.json 2>> gpurun_out/bench2.err
timeout 300 python bench.py --steps 100 --warmup 10 --rebuild 1 --no-cpu-baseline > gpurun_out/bench2_rebuild1.json 2>> gpurun_out/bench2.err
timeout 300 python bench.py --steps 200 --warmup 20 --skin 0.25 --no-cpu-baseline > gpurun_out/bench2_list_skin025.json 2>> gpurun_out/bench2.err
timeout 300 python profiles/prof_step.py lj4m 12 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches2.csv python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_launches2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:az_k_pair -s 1 -c 3 -o gpurun_out/prof_pair2 python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_full2.log 2>&1
echo done
