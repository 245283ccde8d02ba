set -x
mkdir -p gpurun_out
# quad-row list layout (new lib) against the previous build (profiles/r2/ab/libaztot_b200_old.so), same job
B="--steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2"
timeout 300 python bench.py $B > gpurun_out/r2_45_bench_quad_a.json 2> gpurun_out/r2_45_bench_quad_a.err; echo "rc=$?" >> gpurun_out/r2_45_bench_quad_a.err
AZTOT_B200_LIB=$PWD/profiles/r2/ab/libaztot_b200_old.so timeout 300 python bench.py $B > gpurun_out/r2_45_bench_old_a.json 2> gpurun_out/r2_45_bench_old_a.err; echo "rc=$?" >> gpurun_out/r2_45_bench_old_a.err
timeout 300 python bench.py $B > gpurun_out/r2_45_bench_quad_b.json 2> gpurun_out/r2_45_bench_quad_b.err; echo "rc=$?" >> gpurun_out/r2_45_bench_quad_b.err
AZTOT_B200_LIB=$PWD/profiles/r2/ab/libaztot_b200_old.so timeout 300 python bench.py $B > gpurun_out/r2_45_bench_old_b.json 2> gpurun_out/r2_45_bench_old_b.err; echo "rc=$?" >> gpurun_out/r2_45_bench_old_b.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_45_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_45_pytest_gpu_all.log
echo done
