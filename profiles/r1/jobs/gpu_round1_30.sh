mkdir -p gpurun_out
L=aztotmd_b200/lib
for v in minb1 minb6 minb8; do
  cp $L/variants/$v/libaztot_b200.so $L/libaztot_b200.so
  for w in cs1 lj32k lj4k; do
  timeout 200 python bench.py --workload $w --rebuild 1 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 --graph > gpurun_out/bench34_${w}_$v.json 2> gpurun_out/bench34_${w}_$v.err
  python -c "import json; d=json.loads(open('gpurun_out/bench34_${w}_$v.json').read()); print('$w $v', '%.4g'%d['value'], round(d['ms_per_step'],5), round(d['kernel_ms_per_step']['pair'],4))" >> gpurun_out/bench34_summary.txt 2>&1
  done
done
