mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests12.log
for k in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench16_$k.json 2> gpurun_out/bench16_$k.err; echo "rc=$?" >> gpurun_out/bench16_$k.err
done
timeout 300 python bench.py --workload nacl1m --rebuild 5 --skin 0.6 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench16_nacl1m.json 2> gpurun_out/bench16_nacl1m.err
