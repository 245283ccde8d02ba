set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/bench1.err
timeout 300 python bench.py --steps 100 --warmup 10 --rebuild 1 --no-cpu-baseline > gpurun_out/bench1_rebuild1.json 2>> gpurun_out/bench1.err
timeout 300 python profiles/prof_step.py lj4m 12 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:az_k_pair -s 2 -c 2 -o gpurun_out/prof_pair python profiles/prof_step.py lj4m 12 > gpurun_out/ncu_full.log 2>&1
echo done
