set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q -x > gpurun_out/gpu_all3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_all3.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke3.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "rc=$?" >> gpurun_out/bench17.err
