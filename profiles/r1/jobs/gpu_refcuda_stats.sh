mkdir -p gpurun_out
timeout 600 python profiles/compare_reference_cuda_stats.py gpurun_out/reference_cuda_stats.json > gpurun_out/refcuda_stats.log 2>&1; echo "rc=$?" >> gpurun_out/refcuda_stats.log
