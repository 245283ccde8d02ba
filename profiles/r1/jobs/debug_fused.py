import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import azt
from aztotmd_b200 import abi
from aztotmd_b200.engine import Engine
lib = abi.load_library()
systems = [("lj800", azt.lj_system(ncell=(6, 6, 6), seed=31, temperature=800.0), {}),
           ("lj_list", azt.lj_system(ncell=(6, 6, 6), seed=32, cell_size=9.3), dict(rebuild_every=4, skin=0.8)),
           ("surk", azt.surk_system(n_side=6, n_equil=4, eq_freq=2), {}),
           ("lj_radi_list", azt.lj_system(ncell=(5, 5, 5), seed=33, vel=False, tstat=abi.TSTAT_RADI, temperature=298.0), dict(rebuild_every=3, skin=0.6))]
systems[3][1].species[0].radA, systems[3][1].species[0].radB, systems[3][1].species[0].mxEng = 2.73, 4.731, 0.2
def run(s, kw, mode, n=11, energy=False):
    e = Engine(s, lib=lib, **kw)
    if energy: e.set_option("energy_every_step", 1)
    if mode == "one": e.step(0, n)
    elif mode == "nofuse": e.set_option("fuse_integrate", 0); e.step(0, n)
    else:
        for k in range(n): e.step(k, 1)
    d = e.dump(); e.close(); return d
for name, s, kw in systems:
    for n in (1, 2, 3, 11):
        a = run(s, kw, "one", n); a2 = run(s, kw, "one", n); b = run(s, kw, "single", n); c = run(s, kw, "nofuse", n); ae = run(s, kw, "one", n, energy=True)
        def diff(x, y): return {k: int((x[k] != y[k]).any(-1).sum()) if x[k].ndim > 1 else int((x[k] != y[k]).sum()) for k in ("pos", "vel", "frc")}
        print(name, n, "a-a2", diff(a, a2), "a-b", diff(a, b), "a-nofuse", diff(a, c), "b-nofuse", diff(b, c), "a_energy-b", diff(ae, b), "maxrel", float(np.abs(a["frc"] - b["frc"]).max() / np.abs(b["frc"]).max()), flush=True)
