mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests19.log
for w in cs1 cs2 lj32k; do
timeout 200 python bench.py --workload $w --rebuild 1 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench28_${w}.json 2> gpurun_out/bench28_${w}.err; echo "rc=$?" >> gpurun_out/bench28_${w}.err
python -c "import json; d=json.loads(open('gpurun_out/bench28_${w}.json').read()); print('$w', d['value'], d['ms_per_step'], {k: round(x,4) for k,x in d['kernel_ms_per_step'].items()}, d['gpu_launches'])" >> gpurun_out/bench28_summary.txt 2>&1
done
