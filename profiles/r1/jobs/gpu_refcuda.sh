mkdir -p gpurun_out
timeout 1500 python profiles/time_reference_cuda.py gpurun_out/reference_cuda_vs_engine.json > gpurun_out/refcuda.log 2>&1; echo "rc=$?" >> gpurun_out/refcuda.log
