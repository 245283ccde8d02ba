set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_07_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_z_bricks.py -m gpu -q -x -s > gpurun_out/r2_07_pytest_bricks_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_07_pytest_bricks_2gpu.log
for mode in peer sendrecv; do
  if [ $mode = sendrecv ]; then export AZTOT_NO_PEER_HALO=1; else unset AZTOT_NO_PEER_HALO; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r2_07_bench_n2_$mode.json 2> gpurun_out/r2_07_bench_n2_$mode.err; echo "rc=$?" >> gpurun_out/r2_07_bench_n2_$mode.err
done
unset AZTOT_NO_PEER_HALO
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 20 --no-halo-overlap --no-halo-check > gpurun_out/r2_07_bench_n2_peer_nooverlap.json 2> gpurun_out/r2_07_bench_n2_peer_nooverlap.err
timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-small-deck --e2e-steps 2 > gpurun_out/r2_07_bench_n1.json 2> gpurun_out/r2_07_bench_n1.err
echo done
