mkdir -p gpurun_out
L=aztotmd_b200/lib
cp $L/libaztot_b200.so $L/base.so
for v in base idx1 idx2 idx2pos1 tpb64 tpb256 tpb64idx2 base; do
  if [ $v = base ]; then cp $L/base.so $L/libaztot_b200.so; else cp $L/variants/$v/libaztot_b200.so $L/libaztot_b200.so; fi
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench22_$v.json 2> gpurun_out/bench22_$v.err; echo "rc=$?" >> gpurun_out/bench22_$v.err
  python -c "import json; d=json.loads(open('gpurun_out/bench22_$v.json').read()); print('$v', d['value'], d['ms_per_step'], d['kernel_ms_per_step']['pair_list'], d['kernel_ms_per_step']['pair'], d['kernel_ms_per_step']['integrate_b'], d['energy'])" >> gpurun_out/bench22_summary.txt
done
cp $L/base.so $L/libaztot_b200.so
