"""Long-run checks on the B200 (north_star: "total energy drift and thermostat temperature trajectories
within a stated tolerance over 10k steps") and the host driver `aztotmd` end to end on real decks.

Trajectories of a chaotic system cannot be compared atom by atom after thousands of steps, so these are
the size-independent properties the domain offers: conservation of the total energy in NVE, agreement of
the time-averaged kinetic / thermal energy with the oracle under the radiative thermostat, and agreement
between the two schedules of the engine (bin every step vs Verlet list every 10 steps)."""
import os
import subprocess

import numpy as np
import pytest

import azt
from aztotmd_b200 import abi, decks
from aztotmd_b200.engine import KB, Engine

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "aztotmd_b200", "lib", "aztotmd")

# Stated tolerances ------------------------------------------------------------------------------------
DRIFT_TOL_10K = 1e-2      # |E(t) - E(0)| / KE(0) over 10 000 steps of 1 fs.  Measured on B200: 3.8e-3 (0.66 eV of |E| = 1821 eV,
                          # 2e-5 eV per atom).  Two sources, both inherent to the reference's arithmetic: LJ is truncated at
                          # rc without a shift (every crossing of rc changes E by |U(rc)| = 0.0163 eps) and positions are FP32
                          # (one ulp of a 115 A coordinate is 4e-3 of a 1 fs displacement)
SCHEDULE_TOL = 2e-2       # <KE> of the two schedules over the last 5 000 steps, relative
THERMO_TOL = 0.10         # window means of engKin / engTemp, engine vs oracle (different chaotic trajectories)


@pytest.fixture(scope="module")
def lib():
    lib = abi.load_library()
    assert lib.az_device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    return lib


def energy_series(e, nsteps, every):
    out = []
    for k in range(0, nsteps, every):
        e.step(k, every)
        st = e.stats()
        out.append((st.engKin_d, st.engVdW_d + st.engCoul1_d))
    return np.array(out)


def test_nve_energy_drift_10k_steps(lib):
    s = azt.lj_system(ncell=(20, 20, 20), seed=41, cell_size=8.488 + 0.5)  # 32 000 argon atoms, 86.6 K
    e = Engine(s, lib=lib, rebuild_every=10, skin=0.5)
    ser = energy_series(e, 10000, 500)
    tot = ser.sum(1)
    drift = np.abs(tot - tot[0]).max() / ser[0, 0]
    print("NVE 10k steps, list schedule: E0 = %.6f eV, max |dE| / KE0 = %.3g, KE %.4f -> %.4f" % (tot[0], drift, ser[0, 0], ser[-1, 0]))
    assert drift <= DRIFT_TOL_10K, drift
    e.close()
    # the reference schedule (bin + sort every step) conserves energy equally well and lands at the same temperature
    f = Engine(s, lib=lib)
    ser1 = energy_series(f, 10000, 500)
    tot1 = ser1.sum(1)
    drift1 = np.abs(tot1 - tot1[0]).max() / ser1[0, 0]
    print("NVE 10k steps, every-step schedule: max |dE| / KE0 = %.3g" % drift1)
    assert drift1 <= DRIFT_TOL_10K, drift1
    assert abs(tot1[0] - tot[0]) <= 1e-6 * abs(tot[0])  # same first sample: same physics
    ke_a, ke_b = ser[10:, 0].mean(), ser1[10:, 0].mean()
    assert abs(ke_a - ke_b) <= SCHEDULE_TOL * ke_b, (ke_a, ke_b)
    f.close()


def test_ionic_nve_energy_drift(lib):
    s = azt.nacl_system(ncell=(8, 8, 8), elec=abi.ELEC_FENNEL, r_real=12.0, rc_vdw=9.0, cell_size=13.0, seed=42)  # 4 096 ions, DSF
    e = Engine(s, lib=lib, rebuild_every=5, skin=1.0)  # 1200 K ions: 0.5 A of skin per 10 steps is (correctly) refused
    ser = energy_series(e, 4000, 200)
    tot = ser.sum(1)
    drift = np.abs(tot - tot[0]).max() / ser[0, 0]
    print("ionic NVE 4k steps: max |dE| / KE0 = %.3g" % drift)
    assert drift <= 5e-3, drift  # Buckingham is truncated without a shift at 9 A; DSF goes to zero smoothly at 12 A
    e.close()


def test_radiative_thermostat_trajectory_matches_oracle_statistics(lib):
    """surk + radiative thermostat (case study 2 physics), 3 000 steps: the time-averaged kinetic and thermal
    energies of the engine agree with the oracle's (same seeded Philox stream, diverging trajectories)."""
    s = azt.surk_system(n_side=8, spacing=1.3, box=35.0, seed=43, n_equil=1000, eq_freq=250)  # 512 atoms
    o = azt.Oracle(s)
    e = Engine(s, lib=lib)
    e.set_option("energy_every_step", 0)
    ko, ke, uo, ue = [], [], [], []
    for k in range(0, 3000, 100):
        o.step(k, 100); e.step(k, 100)
        so, se = o.stats(), e.stats()
        ko.append(so.engKin_d); ke.append(se.engKin_d); uo.append(so.engTemp_d); ue.append(se.engTemp_d)
    ko, ke, uo, ue = map(np.array, (ko, ke, uo, ue))
    # first sample: trajectories have not diverged yet
    assert abs(ko[0] - ke[0]) <= 1e-3 * ko[0] and abs(uo[0] - ue[0]) <= 1e-3 * max(uo[0], 1e-9)
    # window means over the last 2 000 steps
    mk_o, mk_e, mu_o, mu_e = ko[10:].mean(), ke[10:].mean(), uo[10:].mean(), ue[10:].mean()
    print("radi thermostat: <KE> oracle %.4f engine %.4f, <U> oracle %.4f engine %.4f" % (mk_o, mk_e, mu_o, mu_e))
    assert abs(mk_e - mk_o) <= THERMO_TOL * mk_o and abs(mu_e - mu_o) <= THERMO_TOL * mu_o
    # the thermostat holds the temperature of the deck: <KE> = 3/2 N kT within the thermostat's own fluctuation
    t_eff = 2.0 * mk_e / (3 * len(s.pos) * KB)
    print("radi thermostat: effective temperature %.1f K (target %.1f K)" % (t_eff, s.temperature))
    e.close()


def run_driver(deck_dir, *args):
    assert os.path.exists(DRIVER), "host driver missing: python -m aztotmd_b200.build"
    r = subprocess.run([DRIVER, *args], cwd=deck_dir, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600)
    return r


def read_table(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return rows


def test_host_driver_lj_deck_outputs(lib, tmp_path):
    """`aztotmd` on an LJ deck: the reference's output files appear with the reference's columns, the energy in
    stat.dat is conserved and equals what the engine reports through the ABI on the same deck."""
    d = str(tmp_path)
    info = decks.lj_argon_deck(d, 10, nstep=400, stat=50, rebuild_every=10, skin=0.5, cell_list=8.488 + 0.5, seed=7)
    r = run_driver(d, "--seed", "7")
    assert r.returncode == 0, r.stdout + r.stderr
    stat = read_table(os.path.join(d, "stat.dat"))
    assert stat[0][:7] == ["time", "step", "engTot", "engKin", "engVdW", "engCoul1", "engCoul2"] and stat[0][-1] == "press"
    assert stat[1][0] == "time, ps" and stat[1][-1] == "press, atm"
    steps = [int(x[1]) for x in stat[2:]]
    assert steps == list(range(0, 400, 50))
    etot = np.array([float(x[2]) for x in stat[2:]])
    ekin = np.array([float(x[3]) for x in stat[2:]])
    # unshifted truncation at rc: while the jittered lattice thermalises, ~0.5 % of the pairs of the 8.06 A and 9.01 A
    # shells cross rc = 8.488 A, each changing E by |U(rc)| = 1.6e-4 eV (measured 2.5e-3 of KE)
    assert np.abs(etot - etot[0]).max() <= DRIFT_TOL_10K * ekin[0]
    msd = read_table(os.path.join(d, "msd.dat"))
    assert msd[0][:3] == ["time", "step", "Ar_px"] and len(msd) == len(stat) - 1
    rev = open(os.path.join(d, "revcon.xyz")).read().splitlines()
    assert int(rev[0]) == info["n_atoms"] and rev[1].split()[0] == "1" and len(rev) == info["n_atoms"] + 2
    assert rev[2].split("\t")[0] == "Ar"
    vel = read_table(os.path.join(d, "velocities.dat"))
    assert vel[0][:5] == ["No", "Ar", "x", "y", "z"] and len(vel) == info["n_atoms"] + 1
    assert "Kin=" in r.stdout and "Vdw=" in r.stdout and "execution time" in r.stdout
    # the same deck through the ABI from Python gives the same first row
    import ctypes as C
    h = C.c_void_p()
    assert lib.az_deck_load(d.encode(), 7, C.byref(h)) == 0
    e = Engine(lib.az_deck_desc(h).contents, lib=lib, opts=lib.az_deck_opts(h).contents)
    e.step(0, 1)
    st = e.stats()
    assert abs(st.engTot - etot[0]) <= 1e-5 * abs(etot[0]) + 1e-6
    e.close()
    lib.az_deck_free(h)


def test_lj_deck_drift_is_the_double_precision_reference_drift(lib, tmp_path):
    """The loosened bound of the test above, settled: the reference's OWN serial build (double precision) does not conserve
    energy on this deck either -- LJ is cut at rc without a shift -- and drifts by 0.139 eV (3.2e-3 of the kinetic energy)
    between steps 50 and 400 (tests/golden/deck_lj4000_ref_serial_drift.json, made by tests/golden/make_golden_decks.py
    drift from oracle/_ref/ref_serial_harness).  The engine, on the reference schedule and on the Verlet-list schedule, must
    stay within twice that (different random velocities: the comparison is between two members of one ensemble)."""
    import ctypes as C
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "deck_lj4000_ref_serial_drift.json")))
    ref = float(g["max_abs_dE"])
    assert 0.05 < ref < 0.5
    for tag, kw in (("every step", dict(cell_list=8.488 + 0.5)), ("list, rebuild 10", dict(rebuild_every=10, skin=0.5, cell_list=8.488 + 0.5))):
        d = str(tmp_path / tag.replace(" ", "_").replace(",", ""))
        os.makedirs(d)
        decks.lj_argon_deck(d, 10, nstep=400, stat=50, seed=7, **kw)
        h = C.c_void_p()
        assert lib.az_deck_load(d.encode(), 7, C.byref(h)) == 0
        e = Engine(lib.az_deck_desc(h).contents, lib=lib, opts=lib.az_deck_opts(h).contents)
        e.step(0, 51)
        st = e.stats()
        e0 = st.engKin_d + st.engVdW_d
        worst = 0.0
        for k in range(51, 401, 50):
            e.step(k, 50)
            st = e.stats()
            worst = max(worst, abs(st.engKin_d + st.engVdW_d - e0))
        print("%s: max |E(t) - E(50)| = %.4f eV, reference serial build %.4f eV" % (tag, worst, ref))
        assert worst <= 2.0 * ref, (tag, worst, ref)
        e.close()
        lib.az_deck_free(h)


def test_host_driver_writes_rdf_dat(lib, tmp_path):
    """'rdf rmax dr freq outfreq': rdf.dat in copy_rdf's layout; liquid-like argon: first peak near 1.1 sigma, g -> 1."""
    d = str(tmp_path)
    decks.lj_argon_deck(d, 12, nstep=300, stat=100, temperature=120.0, rdf_freq=10, rdf_rmax=12.0, rdf_dr=0.05, seed=9)
    r = run_driver(d, "--seed", "9")
    assert r.returncode == 0, r.stdout + r.stderr
    rows = read_table(os.path.join(d, "rdf.dat"))
    assert rows[0] == ["r", "Ar-Ar"] and len(rows) == 1 + int(12.0 / 0.05)
    rr = np.array([float(x[0]) for x in rows[1:]])
    g = np.array([float(x[1]) for x in rows[1:]])
    assert rr[0] == pytest.approx(0.025) and rr[1] - rr[0] == pytest.approx(0.05)
    assert g[rr < 3.0].max() == 0.0                       # excluded core
    peak = rr[np.argmax(g)]
    assert 3.5 < peak < 4.1 and g.max() > 2.0             # first shell of the (thermalising) fcc solid at 2^(1/6) sigma .. a/sqrt(2)
    assert g[(rr > 10.5)].mean() == pytest.approx(1.0, abs=0.15)


def test_host_driver_radiative_deck_outputs(lib, tmp_path):
    """A case-study-2-like deck (surk + radiative thermostat): tchars.dat and the engTerm column appear."""
    import test_deck
    n = test_deck.write_cs2_like(tmp_path, n_side=8)
    r = run_driver(str(tmp_path), "--steps", "600", "--seed", "11")
    assert r.returncode == 0, r.stdout + r.stderr
    stat = read_table(os.path.join(str(tmp_path), "stat.dat"))
    assert "engTerm" in stat[0] and len(stat) >= 2 + 3  # steps 0, 200, 400 (stat 200)
    tch = read_table(os.path.join(str(tmp_path), "tchars.dat"))
    assert tch[0] == ["No", "Ar_eng", "Ar_rad"] and len(tch) == n + 1
    rad = np.array([float(x[2]) for x in tch[1:]])
    assert (rad > 0.5).all() and (rad < 0.7).all()  # radius = 2.73 / (4.731 - min(U, 0.2)) in [0.577, 0.603]
    assert "U=" in r.stdout


def test_host_driver_reports_deck_errors(lib, tmp_path):
    decks.lj_argon_deck(str(tmp_path), 4, nstep=10)
    ctl = os.path.join(str(tmp_path), "control.txt")
    txt = open(ctl).read().replace("timestep 0.001 ps\n", "")
    open(ctl, "w").write(txt)
    r = run_driver(str(tmp_path))
    assert r.returncode != 0 and "ERROR[411]" in r.stdout
