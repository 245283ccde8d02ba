"""world_size-2 test of the N > 1 path on the CPU with torch.distributed's gloo backend: two real processes
(torchrun), one emulator-build engine each, halo exchange / migration / all-reduce carried by gloo
send/recv/all_reduce (tests/gloo_brick_worker.py).  The production transport is NCCL (tests/test_gpu_z_bricks.py)."""
import os
import subprocess
import sys

import azt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_over_gloo():
    azt.emu_lib()  # build once, before two processes race for it
    env = dict(os.environ, OMP_NUM_THREADS="1", AZ_EMU_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "tests", "gloo_brick_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok:") == 5, r.stdout
