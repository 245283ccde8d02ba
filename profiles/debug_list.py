"""GPU debugging aid: upload (e2e path) after a Verlet-list run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from aztotmd_b200.engine import Engine, AzError

for skin, ncell, nwarm in [(0.25, 50, 30), (0.25, 100, 30), (0.25, 100, 220), (0.4, 100, 220)]:
    s = bench.lj_system(ncell, 10, skin)
    e2 = Engine(s, rebuild_every=10, skin=skin)
    print("skin", skin, "ncell", ncell, "cnum", tuple(e2.info.cnum), "rows", e2.refresh_info().list_rows, flush=True)
    e2.step(0, nwarm)
    inf = e2.refresh_info(); st = e2.stats()
    print("  before: max_neigh %d flags %d pairs %d" % (inf.max_neigh_seen, inf.flags, st.n_pairs), flush=True)
    a = e2.download()
    d0 = e2.dump()
    print("  download pos range", a["pos"].min(0), a["pos"].max(0), "device pos range", d0["pos"].min(0), d0["pos"].max(0), flush=True)
    for k in range(2):
        e2.upload(a["pos"], a["vel"])
        d1 = e2.dump()
        srt = np.argsort(d1["ids"])
        print("  upload check: pos equal", np.array_equal(d1["pos"][srt].astype(np.float64), a["pos"].astype(np.float32).astype(np.float64)),
              "vel equal", np.array_equal(d1["vel"][srt].astype(np.float64), a["vel"].astype(np.float32).astype(np.float64)), "frc zero", not d1["frc"].any(), flush=True)
        e2.step(nwarm + k, 1)
        inf = e2.refresh_info()
        d2 = e2.dump()
        print("  after upload+step %d: max_neigh %d flags %d count max %d sorted %s posrange %s %s nan %d" % (
            k, inf.max_neigh_seen, inf.flags, d2["count"].max(), bool(np.all(np.diff(d2["cell"]) >= 0)), d2["pos"].min(0), d2["pos"].max(0), np.isnan(d2["pos"]).sum()), flush=True)
        if inf.flags:
            break
    e2.close()
