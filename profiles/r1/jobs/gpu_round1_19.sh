mkdir -p gpurun_out
for v in 1 2 3; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 --cell-div $v > gpurun_out/bench23_div$v.json 2> gpurun_out/bench23_div$v.err; echo "rc=$?" >> gpurun_out/bench23_div$v.err
  python -c "import json; d=json.loads(open('gpurun_out/bench23_div$v.json').read()); print('div$v', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['energy'])" >> gpurun_out/bench23_summary.txt
done
