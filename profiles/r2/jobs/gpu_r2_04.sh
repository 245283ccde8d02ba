set -x
mkdir -p gpurun_out
timeout 1500 python profiles/ref_cuda_probe.py gpurun_out/r2_04_reference_cuda.json gpurun_out/r2_04_reference_cuda_memcheck.log > gpurun_out/r2_04_refcuda.log 2>&1; echo "rc=$?" >> gpurun_out/r2_04_refcuda.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "autocap or lj_basic or rebuild_interval_list or list_every" > gpurun_out/r2_04_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_04_pytest.log
