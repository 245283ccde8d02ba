set -x
mkdir -p gpurun_out
# list build: register budgets (launch bounds 1 / 10 / 12 blocks = 66 / 48 / 40 registers) and inner-loop unroll 1 / 3 / 4 against the default (56 registers, unroll 2), same job
B="--steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2"
for v in default minb1 minb10 minb12 unroll1 unroll3 unroll4 default2; do
  L=$PWD/aztotmd_b200/lib/variants/$v/libaztot_b200.so
  [ $v = default ] && L=$PWD/aztotmd_b200/lib/libaztot_b200.so
  [ $v = default2 ] && L=$PWD/aztotmd_b200/lib/libaztot_b200.so
  AZTOT_B200_LIB=$L timeout 300 python bench.py $B > gpurun_out/r2_46_bench_$v.json 2> gpurun_out/r2_46_bench_$v.err; echo "rc=$?" >> gpurun_out/r2_46_bench_$v.err
done
echo done
