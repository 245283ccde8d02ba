mkdir -p gpurun_out
for w in cs1 cs2 lj32k; do
for v in 0 1 2 3; do
timeout 200 python bench.py --workload $w --rebuild 1 --pair-variant $v --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench26_${w}_v$v.json 2> gpurun_out/bench26_${w}_v$v.err; echo "rc=$?" >> gpurun_out/bench26_${w}_v$v.err
python -c "import json; d=json.loads(open('gpurun_out/bench26_${w}_v$v.json').read()); print('$w v$v', d['value'], d['ms_per_step'], {k: round(x,4) for k,x in d['kernel_ms_per_step'].items()}, d['gpu_launches'])" >> gpurun_out/bench26_summary.txt 2>&1
done
done
