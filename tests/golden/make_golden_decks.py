"""Golden outputs for the reference's two SHIPPED decks (tests/golden/decks/cs1, cs2 = byte copies of
/root/reference/case study 1|2).  Run in the development container; the fixtures are committed.

cs1 (40 000 argon atoms, LJ rc 4 A, dilute gas): the reference's OWN serial build
    (oracle/_ref/ref_serial_harness, compiled from /root/reference/src by oracle/build_ref.sh) runs the deck with
    `temperature 298 none` (the serial twin has no radiative thermostat, BASELINE.json configs[0]) and
    `init_vel gaus` under srand(1); the fixture keeps the velocities it drew, the shift center_box applied, and its
    energies / forces / positions after NSTEP steps -> tests/golden/deck_cs1_ref_serial.npz
cs2 (4 000 atoms, surk + radiative thermostat, 10 000 steps): no build of the reference computes this deck here (the
    serial twin lacks surk and the radiative thermostat, the CUDA twin does not run on sm_100a), so the fixture is the
    CPU restatement's (oracle/az_oracle.c, pinned on the reference's device source) window statistics over 10 000
    steps -> tests/golden/deck_cs2_oracle_windows.json

usage: python tests/golden/make_golden_decks.py [cs1] [cs2] [drift]
"""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
from make_golden import HARNESS, read_dump  # noqa: E402

CS1_NSTEP = 20
CS2_NSTEP, CS2_EVERY = 10000, 250


def deck_positions(d):
    rows = open(os.path.join(d, "atoms.xyz")).read().splitlines()
    n = int(rows[0])
    return np.array([[float(x) for x in r.split()[1:4]] for r in rows[2:2 + n]])


def make_cs1():
    src = os.path.join(HERE, "decks", "cs1")
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "cs1")
        shutil.copytree(src, d)
        ctl = open(os.path.join(d, "control.txt"), encoding="latin-1").read()
        ctl = ctl.replace("temperature 298.0\tradi\t0.2", "temperature 298.0\tnone").replace("init_vel\tzero\t0.0332", "init_vel\tgaus")
        assert "temperature 298.0\tnone" in ctl and "init_vel\tgaus" in ctl
        open(os.path.join(d, "control.txt"), "w", encoding="latin-1").write(ctl + "\nhist 100000\n")  # hist 0 is a SIGFPE in the serial main
        dump = os.path.join(d, "dump.bin")
        r = subprocess.run([HARNESS, str(CS1_NSTEP), "fast", dump, "1"], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True, check=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        dumps = read_dump(dump)
    init, fin = dumps["init"], dumps["final"]
    p0 = deck_positions(src)
    L = 1141.5
    dd = init["pos"] - p0
    shift = np.median(dd, 0)
    wrapped = dd - shift - L * np.round((dd - shift) / L)   # integrate1_clst wraps the shifted atoms back into the box
    assert np.abs(wrapped).max() < 1e-9, "center_box is a rigid shift (modulo the box)"
    nz = np.nonzero(np.abs(fin["frc"]).max(1) > 0)[0]
    out = os.path.join(HERE, "deck_cs1_ref_serial.npz")
    np.savez_compressed(out, shift=shift, pos0_sub=init["pos"][::16], vel0=init["vel"], nstep=CS1_NSTEP, engKin0=init["engKin"], engVdW0=init["engVdW"],
                        engKin=fin["engKin"], engVdW=fin["engVdW"], frc_idx=nz.astype(np.int32), frc=fin["frc"][nz],
                        pos_sub=fin["pos"][::16], vel_sub=fin["vel"][::16], harness=np.array(line))
    print("cs1", line)
    print("  pairs with force after %d steps: %d atoms; engVdW %.6e, engKin %.6f" % (CS1_NSTEP, len(nz), fin["engVdW"], fin["engKin"]))
    print("  ->", out, os.path.getsize(out), "bytes")


def make_drift():
    """The NVE energy series of the reference's serial build (double precision) on the 4 000-atom LJ deck of
    tests/test_gpu_long.py::test_host_driver_lj_deck_outputs (decks.lj_argon_deck(10, seed 7), rc without a shift): the
    energy is not conserved there either -- every pair that crosses rc changes E by |U(rc)| -- and the engine's drift
    (from step 50 on, like this series) is judged against this one, not against zero -> tests/golden/deck_lj4000_ref_serial_drift.json"""
    from aztotmd_b200 import decks
    from make_golden import run
    steps = list(range(0, 401, 50))
    series = []
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "lj4000")
        os.makedirs(d)
        info = decks.lj_argon_deck(d, 10, nstep=400, stat=50, cell_list=8.488 + 0.5, seed=7)
        open(os.path.join(d, "control.txt"), "a").write("\nhist 100000\n")
        for n in steps[1:]:
            dumps, line = run(d, n, "fast")
            series.append((n, dumps["final"]["engKin"], dumps["final"]["engVdW"]))  # (the harness has no kinetic energy before the first step)
    e = np.array([k + v for _, k, v in series])
    out = os.path.join(HERE, "deck_lj4000_ref_serial_drift.json")
    json.dump({"deck": "decks.lj_argon_deck(ncell=10, nstep=400, stat=50, cell_list=8.988, seed=7): 4000 Ar, rc = 2.5 sigma unshifted, dt 1 fs, init_vel gaus under srand(1)",
               "steps": [s for s, _, _ in series], "engKin": [k for _, k, _ in series], "engVdW": [v for _, _, v in series],
               "max_abs_dE": float(np.abs(e - e[0]).max()), "max_abs_dE_over_KE0": float(np.abs(e - e[0]).max() / series[0][1]),
               "source": "oracle/_ref/ref_serial_harness (the reference's own serial sources, double precision)"}, open(out, "w"), indent=1)
    print("drift", [round(float(x - e[0]), 6) for x in e], "max |dE| / KE0 = %.3g" % (np.abs(e - e[0]).max() / series[0][1]))
    print("  ->", out)


def make_cs2():
    import azt
    from aztotmd_b200 import abi
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.az_deck_load(os.path.join(HERE, "decks", "cs2").encode(), 20261018, C.byref(h)) == 0
    d = lib.az_deck_desc(h).contents
    o = azt.Oracle(d)
    rows = []
    for k in range(0, CS2_NSTEP, CS2_EVERY):
        o.step(k, CS2_EVERY)
        st = o.stats()
        rows.append(dict(step=k + CS2_EVERY - 1, engKin=st.engKin_d, engVdW=st.engVdW_d, engTemp=st.engTemp_d))
        print(rows[-1], flush=True)
    out = os.path.join(HERE, "deck_cs2_oracle_windows.json")
    json.dump(dict(deck="tests/golden/decks/cs2", seed=20261018, nstep=CS2_NSTEP, every=CS2_EVERY, n_atoms=int(d.n_atoms),
                   t_kin=float(d.t_kin), n_equil=int(d.n_equil), eq_freq=int(d.eq_freq), rows=rows,
                   source="oracle/az_oracle.c (azo_step), samples of the last step of each window"), open(out, "w"), indent=1)
    lib.az_deck_free(h)
    print("  ->", out)


def make_tstat():
    """a32: what the reference's read_tstat (src/temperature.cpp:91-259) builds for `temperature 298 radi 1` under srand(1):
    the photon energies of photon_engs and the 3072 preset unit vectors -> tests/golden/ref_tstat_tables.npz"""
    from aztotmd_b200 import decks
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "t")
        decks.lj_argon_deck(d, 6, nstep=1, init_vel="zero", tstat="radi\t1", temperature=298.0)
        r = subprocess.run([HARNESS, "0", "fast", "-", "1", "tstat"], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True, check=True)
        raw = open(os.path.join(d, "tstat.bin"), "rb").read()
    n, nv = np.frombuffer(raw[:8], dtype=np.int32)
    a = np.frombuffer(raw[8:], dtype=np.float64)
    photons, uv = a[:n], a[n:].reshape(3, nv)
    assert nv == 3072 and len(a) == n + 3 * nv
    out = os.path.join(HERE, "ref_tstat_tables.npz")
    np.savez_compressed(out, photons=photons, uvect=uv, temperature=298.0, srand=1,
                        source=np.array("oracle/_ref/ref_serial_harness 0 fast - 1 tstat (read_tstat of /root/reference/src/temperature.cpp)"))
    print("tstat: %d photon energies (mean %.5f eV), %d unit vectors -> %s (%d bytes)" % (n, photons.mean(), nv, out, os.path.getsize(out)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["cs1", "cs2", "tstat"]
    if "tstat" in what:
        make_tstat()
    if "cs1" in what:
        make_cs1()
    if "cs2" in what:
        make_cs2()
    if "drift" in what:
        make_drift()
