set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload nacl1m --rebuild 5 --skin 0.6 --steps 50 --warmup 10 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_36_bench_nacl1m.json 2> gpurun_out/r2_36_bench_nacl1m.err; echo "rc=$?" >> gpurun_out/r2_36_bench_nacl1m.err
timeout 600 python bench.py --workload nacl1m_pme --rebuild 5 --skin 0.6 --steps 50 --warmup 10 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_36_bench_nacl1m_pme.json 2> gpurun_out/r2_36_bench_nacl1m_pme.err; echo "rc=$?" >> gpurun_out/r2_36_bench_nacl1m_pme.err
timeout 600 python bench.py --workload lj4m --rebuild 1 --steps 20 --warmup 5 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_36_bench_lj4m_r1.json 2> gpurun_out/r2_36_bench_lj4m_r1.err; echo "rc=$?" >> gpurun_out/r2_36_bench_lj4m_r1.err
echo done
