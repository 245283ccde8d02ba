mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests23_all_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests20_all_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/bench36_n2.json 2> gpurun_out/bench36_n2.err; echo "rc=$?" >> gpurun_out/bench36_n2.err
