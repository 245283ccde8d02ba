set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests6.log
for p in 0 2 4; do
timeout 300 python bench.py --steps 200 --warmup 20 --list-pipe $p --no-cpu-baseline > gpurun_out/bench5_pipe$p.json 2> gpurun_out/bench5_pipe$p.err; echo "rc=$?" >> gpurun_out/bench5_pipe$p.err
done
timeout 300 python bench.py --steps 200 --warmup 20 --list-pipe 0 --cell-div 2 --no-cpu-baseline > gpurun_out/bench5_div2.json 2> gpurun_out/bench5_div2.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/prof_plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"az_k_pair_list|az_k_build_list" -s 0 -c 2 -o gpurun_out/prof_list5 python profiles/prof_step.py lj4m 12 10 0.25 0 0 > gpurun_out/ncu_full5.log 2>&1
echo done
