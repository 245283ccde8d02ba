mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests10.log
for w in cs1 cs2 lj32k; do
timeout 200 python bench.py --workload $w --rebuild 1 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err; echo "rc=$?" >> gpurun_out/bench2_$w.err
done
timeout 200 python bench.py --rebuild 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench2_lj4m_r1.json 2> gpurun_out/bench2_lj4m_r1.err
timeout 200 python bench.py --rebuild 1 --steps 50 --warmup 5 --pair-variant 1 --no-cpu-baseline > gpurun_out/bench2_lj4m_r1_tiles.json 2> gpurun_out/bench2_lj4m_r1_tiles.err
