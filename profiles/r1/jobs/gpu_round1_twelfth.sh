set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q -x > gpurun_out/gpu_all2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_all2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke2.log
timeout 400 python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "rc=$?" >> gpurun_out/bench15.err
