mkdir -p gpurun_out
for d in 0 1 2; do
timeout 300 python bench.py --steps 100 --warmup 10 --list-dbg $d --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench6_dbg$d.json 2> gpurun_out/bench6_dbg$d.err; echo "rc=$?" >> gpurun_out/bench6_dbg$d.err
done
