mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests22.log
for w in cs1 cs2 lj32k lj4k; do
for v in 3 4; do
timeout 200 python bench.py --workload $w --rebuild 1 --pair-variant $v --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 --graph > gpurun_out/bench33_${w}_v$v.json 2> gpurun_out/bench33_${w}_v$v.err; echo "rc=$?" >> gpurun_out/bench33_${w}_v$v.err
python -c "import json; d=json.loads(open('gpurun_out/bench33_${w}_v$v.json').read()); print('$w v$v', '%.4g'%d['value'], round(d['ms_per_step'],5), {k: round(x,4) for k,x in d['kernel_ms_per_step'].items() if x>0})" >> gpurun_out/bench33_summary.txt 2>&1
done
done
