"""Where does the end-to-end step (host buffers in, stats out) spend its time?  usage: python profiles/e2e_probe.py [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aztotmd_b200 import abi  # noqa: E402
from aztotmd_b200.engine import Engine  # noqa: E402

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ncell, _ = bench.WORKLOADS["lj4m"]
s = bench.lj_system(ncell, 10, 0.25)
lib = abi.load_library()
e = Engine(s, rebuild_every=10, skin=0.25)
n = e.n
e.step(0, 20)
a = e.download()
fpt = C.POINTER(C.c_float)
hp = lib.az_host_alloc(6 * 4 * n)
host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(6, n))
host[0:3] = a["pos"].T
host[3:6] = a["vel"].T
ptr = [C.cast(hp + 4 * n * k, fpt) for k in range(6)]
up = lib.az_engine_upload_f32_async
first = 20


def timed(label, fn, reps):
    e.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    e.sync()
    dt = (time.perf_counter() - t0) / reps * 1e3
    print("%-46s %.3f ms per iteration" % (label, dt))
    return dt


def only_upload():
    assert up(e._h, *ptr) == 0


def upload_step():
    global first
    assert up(e._h, *ptr) == 0
    e.step(first, 1); first += 1


def upload_step_stats():
    upload_step(); e.stats()


def pipelined():
    global first
    e.step(first, 1); first += 1
    assert up(e._h, *ptr) == 0
    e.stats()


for _ in range(2):
    upload_step_stats()
timed("upload only (H2D 96 MB + unpack kernel)", only_upload, ne)
timed("upload + step, no stats, back to back", upload_step, ne)
timed("upload + step + stats (serial)", upload_step_stats, ne)
only_upload()
timed("step k | upload k+1 | stats k (bench e2e loop)", pipelined, ne)
e.profile(2)
only_upload()
for _ in range(ne):
    pipelined()
e.sync()
tm = e.timers()
e.profile(0)
print({k: round(v["ms"] / ne, 4) for k, v in tm.items()})
