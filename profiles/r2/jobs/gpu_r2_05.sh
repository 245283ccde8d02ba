set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_05_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_05_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 20 > gpurun_out/r2_05_bench_default.json 2> gpurun_out/r2_05_bench_default.err; echo "rc=$?" >> gpurun_out/r2_05_bench_default.err
AZ_REF_N1=50 AZ_REF_N2=250 timeout 900 python profiles/ref_cuda_probe.py gpurun_out/r2_05_reference_cuda.json gpurun_out/r2_05_memcheck.log > gpurun_out/r2_05_refcuda.log 2>&1; echo "rc=$?" >> gpurun_out/r2_05_refcuda.log
