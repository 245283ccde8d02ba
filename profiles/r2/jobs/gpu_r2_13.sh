set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 profiles/r2/micro/tile_gather_probe2 > gpurun_out/r2_13_probe2.txt 2>&1; echo rc=$?
echo done
