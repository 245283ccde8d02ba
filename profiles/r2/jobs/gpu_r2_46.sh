set -x
mkdir -p gpurun_out
# list build at 48 / 40 registers (launch bounds 10 / 12 blocks) against the default 56, same job
B="--steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2"
for v in default minb10 minb12 default; do
  L=$PWD/aztotmd_b200/lib/libaztot_b200.so
  [ $v != default ] && L=$PWD/aztotmd_b200/lib/variants/$v/libaztot_b200.so
  AZTOT_B200_LIB=$L timeout 300 python bench.py $B > gpurun_out/r2_46_bench_$v.json 2> gpurun_out/r2_46_bench_$v.err; echo "rc=$?" >> gpurun_out/r2_46_bench_$v.err
done
echo done
