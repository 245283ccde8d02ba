set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_28_pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_28_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_28_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_28_smoke.log
timeout 900 python bench.py > gpurun_out/r2_28_bench_default.json 2> gpurun_out/r2_28_bench_default.err; echo "rc=$?" >> gpurun_out/r2_28_bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_28_bench_reference.json 2> gpurun_out/r2_28_bench_reference.err; echo "rc=$?" >> gpurun_out/r2_28_bench_reference.err
timeout 300 python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_28_prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_28_launches.csv python profiles/prof_step.py lj4m 12 10 0.25 0 > gpurun_out/r2_28_ncu_launches.log 2>&1
echo done
