set -x
mkdir -p gpurun_out
timeout 600 python profiles/e2e_probe.py 10 > gpurun_out/r2_19_e2e_probe.txt 2>&1
echo done
