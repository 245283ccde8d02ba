mkdir -p gpurun_out
for w in cs1 cs2 lj32k; do
for g in "" "--graph"; do
timeout 200 python bench.py --workload $w --rebuild 1 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 $g > gpurun_out/bench32_${w}_r1$g.json 2> gpurun_out/bench32_${w}_r1$g.err; echo "rc=$?" >> gpurun_out/bench32_${w}_r1$g.err
python -c "import json; d=json.loads(open('gpurun_out/bench32_${w}_r1$g.json').read()); print('$w r1 $g', d['value'], d['ms_per_step'], d["section_events"]["ms_per_step"], d.get('cuda_graph'), d['gpu_launches'])" >> gpurun_out/bench32_summary.txt 2>&1
done
done
for g in "" "--graph"; do
timeout 200 python bench.py --workload cs1 --rebuild 10 --skin 10 --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 1 $g > gpurun_out/bench32_cs1_r10$g.json 2> gpurun_out/bench32_cs1_r10$g.err
python -c "import json; d=json.loads(open('gpurun_out/bench32_cs1_r10$g.json').read()); print('cs1 r10 $g', d['value'], d['ms_per_step'], d["section_events"]["ms_per_step"], d.get('cuda_graph'))" >> gpurun_out/bench32_summary.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 1 $g > gpurun_out/bench32_lj4m$g.json 2> gpurun_out/bench32_lj4m$g.err
python -c "import json; d=json.loads(open('gpurun_out/bench32_lj4m$g.json').read()); print('lj4m $g', d['value'], d['ms_per_step'], d["section_events"]["ms_per_step"], d.get('cuda_graph'), d['gpu_launches'])" >> gpurun_out/bench32_summary.txt 2>&1
done
timeout 400 python bench.py > gpurun_out/bench32_default.json 2> gpurun_out/bench32_default.err; echo "rc=$?" >> gpurun_out/bench32_default.err
