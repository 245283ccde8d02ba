mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or rebuild or lj_basic or determinism or full_size" > gpurun_out/gpu_tests15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests15.log
for k in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench21_remap_$k.json 2> gpurun_out/bench21_remap_$k.err; echo "rc=$?" >> gpurun_out/bench21_remap_$k.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --option tile_remap=0 > gpurun_out/bench21_plain_$k.json 2> gpurun_out/bench21_plain_$k.err; echo "rc=$?" >> gpurun_out/bench21_plain_$k.err
done
