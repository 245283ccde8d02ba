mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_tests21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests21.log
for w in cs1 cs2 lj32k; do
timeout 200 python bench.py --workload $w --rebuild 1 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 --graph > gpurun_out/bench30_${w}_graph.json 2> gpurun_out/bench30_${w}_graph.err; echo "rc=$?" >> gpurun_out/bench30_${w}_graph.err
python -c "import json; d=json.loads(open('gpurun_out/bench30_${w}_graph.json').read()); print('$w', d['value'], d['ms_per_step'], d['cuda_graph'], d['gpu_launches'])" >> gpurun_out/bench30_summary.txt 2>&1
done
timeout 200 python bench.py --workload cs1 --rebuild 10 --skin 10 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 --graph > gpurun_out/bench30_cs1_r10_graph.json 2> gpurun_out/bench30_cs1_r10_graph.err
python -c "import json; d=json.loads(open('gpurun_out/bench30_cs1_r10_graph.json').read()); print('cs1 r10', d['value'], d['ms_per_step'], d['cuda_graph'])" >> gpurun_out/bench30_summary.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 --graph > gpurun_out/bench30_lj4m_graph.json 2> gpurun_out/bench30_lj4m_graph.err
python -c "import json; d=json.loads(open('gpurun_out/bench30_lj4m_graph.json').read()); print('lj4m', d['value'], d['ms_per_step'], d['cuda_graph'])" >> gpurun_out/bench30_summary.txt 2>&1
