mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ewald or ionic" > gpurun_out/gpu_tests17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests17.log
for w in nacl64k_pme nacl1m_pme; do
timeout 300 python bench.py --workload $w --rebuild 5 --skin 1.0 --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench25_$w.json 2> gpurun_out/bench25_$w.err; echo "rc=$?" >> gpurun_out/bench25_$w.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:az_k_ewald -c 3 -o gpurun_out/ewald_1m_v2 python bench.py --workload nacl1m_pme --rebuild 5 --skin 1.0 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_ewald.log 2>&1
