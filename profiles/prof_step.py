"""Minimal driver for ncu: a few MD steps of a benchmark workload, nothing else.
usage: python profiles/prof_step.py [workload] [steps] [rebuild] [skin]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aztotmd_b200.engine import Engine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "lj4m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rebuild = int(sys.argv[3]) if len(sys.argv) > 3 else 10
skin = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ncell, _ = bench.WORKLOADS[wl]
s = bench.lj_system(ncell, rebuild, skin)
e = Engine(s, rebuild_every=rebuild, skin=skin if rebuild > 1 else 0.0, pair_variant=variant)
e.step(0, steps)
st = e.stats()
print("ok", wl, steps, st.engKin_d, st.engVdW_d, st.n_pairs, e.launch_count)
e.close()
