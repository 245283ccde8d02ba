set -x
mkdir -p gpurun_out
timeout 300 profiles/r2/micro/tex_gather_probe > gpurun_out/r2_34_tex_probe.txt 2>&1; echo rc=$?
echo done
