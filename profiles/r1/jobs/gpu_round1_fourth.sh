set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi4.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests5.log
for p in 0 2 4; do
timeout 300 python bench.py --steps 200 --warmup 20 --skin 0.25 --list-pipe $p --no-cpu-baseline > gpurun_out/bench4_pipe$p.json 2> gpurun_out/bench4_pipe$p.err; echo "rc=$?" >> gpurun_out/bench4_pipe$p.err
done
timeout 300 python bench.py --steps 200 --warmup 20 --skin 0.4 --list-pipe 0 --no-cpu-baseline > gpurun_out/bench4_skin04.json 2> gpurun_out/bench4_skin04.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/prof_plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches4.csv python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/ncu_launches4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list" -s 0 -c 2 -o gpurun_out/prof_list4 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list" -s 1 -c 1 -o gpurun_out/prof_list4_pipe4 python profiles/prof_step.py lj4m 12 10 0.25 0 4 > gpurun_out/ncu_full4b.log 2>&1
echo done
