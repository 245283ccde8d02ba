"""The reference's two SHIPPED decks on the B200 (BASELINE.json configs[0] and configs[1]).

tests/golden/decks/cs1, cs2 are byte copies of `/root/reference/case study 1|2` (input fixtures).
  * cs1 (40 000 argon atoms, LJ rc 4 A): the engine on the deck with `temperature 298 none` and the velocities the
    reference's serial build drew reproduces that build's energies, forces and positions after 20 steps
    (tests/golden/deck_cs1_ref_serial.npz, made by tests/golden/make_golden_decks.py from oracle/_ref/ref_serial_harness);
    the deck as shipped (radiative thermostat) runs through the host driver `aztotmd` and writes the reference's files.
  * cs2 (4 000 atoms, surk + radiative thermostat, nequil 10 000): 10 000 steps; window statistics of the kinetic,
    pair and thermal energies against the CPU restatement's on the same seeded deck
    (tests/golden/deck_cs2_oracle_windows.json), with asserted tolerances and an asserted effective temperature.
"""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from aztotmd_b200 import abi
from aztotmd_b200.engine import KB, Engine

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
DRIVER = os.path.join(ROOT, "aztotmd_b200", "lib", "aztotmd")

# Stated tolerances ------------------------------------------------------------------------------------
CS1_ENERGY_TOL = 2e-6     # kinetic energy after 20 steps, engine (fp32 state) vs the reference's serial build (fp64), relative
CS1_POS_TOL = 1.3e-3      # A: 20 steps x half an ulp of an fp32 coordinate in [1024, 1141.5) A (ulp = 1.2e-4 A), vs the serial build's fp64
CS1_FORCE_TOL = 2e-3      # the one pair inside rc after 20 steps sits on the steep wall: dF/F = 13 dr/r, dr ~ 2 ulp of 1141 A
CS2_FIRST_TOL = 1e-3      # first window (step 249): trajectories have not diverged yet
CS2_WINDOW_TOL = 0.05     # means over the last 5 000 steps, engine vs oracle (diverged chaotic trajectories, same thermostat stream)
CS2_T_BAND = (0.25, 0.36) # effective kinetic temperature / nominal 500 K during the equilibration period: temp_scale rescales
                          # the kinetic energy to 0.25 * teKin for the radiative thermostat (src/cuTemp.cu:90-94) every 2 500
                          # steps and the thermostat feeds it back slowly in between (oracle: 0.25 -> 0.29 within a period)


@pytest.fixture(scope="module")
def lib():
    lib = abi.load_library()
    assert lib.az_device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    return lib


def load_deck(lib, path, seed):
    h = C.c_void_p()
    rc = lib.az_deck_load(str(path).encode(), seed, C.byref(h))
    assert rc == 0, lib.az_deck_error(h)
    return h


def as_ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_cs1_matches_the_reference_serial_build(lib, tmp_path):
    g = np.load(os.path.join(GOLD, "deck_cs1_ref_serial.npz"))
    d = os.path.join(str(tmp_path), "cs1")
    shutil.copytree(os.path.join(GOLD, "decks", "cs1"), d)
    ctl = open(os.path.join(d, "control.txt"), encoding="latin-1").read().replace("temperature 298.0\tradi\t0.2", "temperature 298.0\tnone")
    open(os.path.join(d, "control.txt"), "w", encoding="latin-1").write(ctl)
    h = load_deck(lib, d, 1)
    desc = lib.az_deck_desc(h).contents
    n = desc.n_atoms
    assert n == 40000 and desc.tstat_kind == abi.TSTAT_NONE
    # the serial build centres the box before it starts (center_box, src/box.cpp:337): apply its shift, keep its velocities
    L = float(desc.la)
    pos = np.stack([np.ctypeslib.as_array(p, shape=(n,)) for p in (desc.xs, desc.ys, desc.zs)], 1).copy()
    pos0 = (pos + g["shift"]) % L
    assert np.abs(pos0[::16] - g["pos0_sub"]).max() < 1e-9
    xs, ys, zs = (np.ascontiguousarray(pos0[:, a]) for a in range(3))
    vel0 = g["vel0"]
    vx, vy, vz = (np.ascontiguousarray(vel0[:, a]) for a in range(3))
    desc.xs, desc.ys, desc.zs = as_ptr(xs), as_ptr(ys), as_ptr(zs)
    desc.vxs, desc.vys, desc.vzs = as_ptr(vx), as_ptr(vy), as_ptr(vz)
    e = Engine(desc, lib=lib, opts=lib.az_deck_opts(h).contents)
    assert tuple(e.info.cnum) == (13, 13, 13)              # split_cells: floor(1141.5 / 85) per axis, bit-exact grid
    e.set_option("energy_every_step", 1)
    e.step(0, 1)
    st = e.stats()
    assert st.engVdW_d == 0.0 and st.n_pairs == 0           # no pair inside 4.0 A in the shipped configuration (min distance 4.008 A)
    nstep = int(g["nstep"])
    e.step(1, nstep - 1)
    st = e.stats()
    assert abs(st.engKin_d - float(g["engKin"])) <= CS1_ENERGY_TOL * float(g["engKin"]), (st.engKin_d, float(g["engKin"]))
    a = e.download(abi.ORDER_ORIGINAL)
    dp = a["pos"][::16] - g["pos_sub"]
    dp -= L * np.round(dp / L)
    assert np.abs(dp).max() <= CS1_POS_TOL
    assert np.abs(a["vel"][::16] - g["vel_sub"]).max() <= 1e-5 * np.abs(g["vel_sub"]).max()
    # the pairs that came inside rc: same atoms, same force
    dmp = e.dump()
    f = np.zeros((n, 3)); f[dmp["ids"]] = dmp["frc"]
    nz = np.nonzero(np.abs(f).max(1) > 0)[0]
    assert np.array_equal(nz, g["frc_idx"]), (nz, g["frc_idx"])
    assert np.abs(f[nz] - g["frc"]).max() <= CS1_FORCE_TOL * np.abs(g["frc"]).max()
    assert abs(st.engVdW_d - float(g["engVdW"])) <= CS1_FORCE_TOL * abs(float(g["engVdW"]))
    e.close()
    lib.az_deck_free(h)


def test_cs1_as_shipped_through_the_host_driver(lib, tmp_path):
    """The deck exactly as shipped (radiative thermostat, `rdf ... nucl`): 400 steps of the driver, the reference's files appear."""
    d = os.path.join(str(tmp_path), "cs1")
    shutil.copytree(os.path.join(GOLD, "decks", "cs1"), d)
    r = subprocess.run([DRIVER, "--steps", "400", "--seed", "3"], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    stat = [l.split("\t") for l in open(os.path.join(d, "stat.dat")).read().splitlines()]
    assert stat[0][:7] == ["time", "step", "engTot", "engKin", "engVdW", "engCoul1", "engCoul2"] and "engTerm" in stat[0]
    assert [int(x[1]) for x in stat[2:]] == [0, 200]
    assert float(stat[2][4]) == 0.0                         # engVdW at step 0
    rev = open(os.path.join(d, "revcon.xyz")).read().splitlines()
    assert int(rev[0]) == 40000 and len(rev) == 40002
    tch = open(os.path.join(d, "tchars.dat")).read().splitlines()
    assert tch[0].split("\t") == ["No", "Ar_eng", "Ar_rad"] and len(tch) == 40001
    assert os.path.exists(os.path.join(d, "rdf_n.dat")), "rdf ... nucl: the by-nuclei RDF file of copy_nrdf (src/cuStat.cu:696-777)"


def test_cs2_ten_thousand_steps_against_the_oracle_windows(lib):
    gold = json.load(open(os.path.join(GOLD, "deck_cs2_oracle_windows.json")))
    h = load_deck(lib, os.path.join(GOLD, "decks", "cs2"), gold["seed"])
    desc = lib.az_deck_desc(h).contents
    assert desc.n_atoms == gold["n_atoms"] == 4000 and desc.tstat_kind == abi.TSTAT_RADI
    assert desc.n_equil == 10000 and desc.eq_freq == 2500
    e = Engine(desc, lib=lib, opts=lib.az_deck_opts(h).contents)
    assert tuple(e.info.cnum) == (12, 12, 12)
    every = gold["every"]
    rows = []
    for k in range(0, gold["nstep"], every):
        e.step(k, every)
        st = e.stats()
        rows.append((st.engKin_d, st.engVdW_d, st.engTemp_d))
    rows = np.array(rows)
    ref = np.array([(r["engKin"], r["engVdW"], r["engTemp"]) for r in gold["rows"]])
    # first window: the two trajectories are still the same one
    assert np.all(np.abs(rows[0] - ref[0]) <= CS2_FIRST_TOL * np.abs(ref[0])), (rows[0], ref[0])
    # the rescale steps report engKin := teKin (after_tscale, src/cuTemp.cu:109-113): sample 2 499 + 1 is step 2 500
    half = len(rows) // 2
    me, mo = rows[half:].mean(0), ref[half:].mean(0)
    print("cs2 10k steps: <KE> engine %.3f oracle %.3f, <VdW> %.2f / %.2f, <U> %.3f / %.3f" % (me[0], mo[0], me[1], mo[1], me[2], mo[2]))
    assert np.all(np.abs(me - mo) <= CS2_WINDOW_TOL * np.abs(mo)), (me, mo)
    # effective kinetic temperature of the equilibration period
    t_eff = 2.0 * rows[half:, 0].mean() / (3 * desc.n_atoms * KB)
    t_nom = 500.0
    print("cs2: effective temperature %.1f K = %.3f of the nominal %.0f K" % (t_eff, t_eff / t_nom, t_nom))
    assert CS2_T_BAND[0] * t_nom <= t_eff <= CS2_T_BAND[1] * t_nom, t_eff
    e.close()
    lib.az_deck_free(h)
