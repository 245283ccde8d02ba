set -x
mkdir -p gpurun_out
# probe: atoms of a cell ranked by a 4x4x4 sub-cell Morton code (ties by id) instead of by id alone -- does the list pass gather fewer lines?
B="--no-cpu-baseline --no-small-deck --e2e-steps 2"
for v in default morton default2 morton2; do
  L=$PWD/aztotmd_b200/lib/variants/morton/libaztot_b200.so
  [ $v = default ] && L=$PWD/aztotmd_b200/lib/libaztot_b200.so
  [ $v = default2 ] && L=$PWD/aztotmd_b200/lib/libaztot_b200.so
  S="--steps 100 --warmup 20"
  [ $v = default2 ] && S="--steps 200 --warmup 400"
  [ $v = morton2 ] && S="--steps 200 --warmup 400"
  AZTOT_B200_LIB=$L timeout 300 python bench.py $B $S > gpurun_out/r2_47_bench_$v.json 2> gpurun_out/r2_47_bench_$v.err; echo "rc=$?" >> gpurun_out/r2_47_bench_$v.err
done
echo done
