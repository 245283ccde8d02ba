"""The process surface of the reference without a GPU: the host driver `aztotmd` (az_main.cpp + the deck
parser az_deck.cpp) linked against the emulator build of the engine (tests/emu, test infrastructure) reads the four
deck files from the working directory and writes stat.dat / msd.dat / rdf.dat / revcon.xyz / velocities.dat in the
reference's formats.  The same checks run against the nvcc build in tests/test_gpu_long.py."""
import ctypes as C
import os
import subprocess

import numpy as np

import azt
from aztotmd_b200 import abi, decks
from aztotmd_b200.engine import Engine

DRIVER = os.path.join(azt.EMU_DIR, "_build", "aztotmd_emu")


def run_driver(deck_dir, *args):
    azt.emu_lib()  # builds the emulator library and the driver
    return subprocess.run([DRIVER, *args], cwd=deck_dir, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, AZ_EMU_THREADS="2"))


def read_table(path):
    return [l.rstrip("\n").split("\t") for l in open(path)]


def test_driver_lj_deck_files_and_formats(tmp_path):
    d = str(tmp_path)
    info = decks.lj_argon_deck(d, 5, nstep=20, stat=5, rebuild_every=5, skin=0.5, cell_list=8.488 + 0.5, seed=7,
                               rdf_freq=5, rdf_rmax=8.0, rdf_dr=0.1)
    r = run_driver(d, "--seed", "7")
    assert r.returncode == 0, r.stdout + r.stderr
    stat = read_table(os.path.join(d, "stat.dat"))
    assert stat[0][:7] == ["time", "step", "engTot", "engKin", "engVdW", "engCoul1", "engCoul2"] and stat[0][-1] == "press"
    assert stat[1][0] == "time, ps" and stat[1][-1] == "press, atm"
    assert [int(x[1]) for x in stat[2:]] == [0, 5, 10, 15]
    etot = np.array([float(x[2]) for x in stat[2:]])
    ekin = np.array([float(x[3]) for x in stat[2:]])
    assert np.abs(etot - etot[0]).max() <= 1e-2 * ekin[0]      # NVE, 20 steps of 1 fs
    msd = read_table(os.path.join(d, "msd.dat"))
    assert msd[0][:3] == ["time", "step", "Ar_px"] and len(msd) == len(stat) - 1
    rev = open(os.path.join(d, "revcon.xyz")).read().splitlines()
    assert int(rev[0]) == info["n_atoms"] and rev[1].split()[0] == "1" and len(rev) == info["n_atoms"] + 2
    vel = read_table(os.path.join(d, "velocities.dat"))
    assert vel[0][:5] == ["No", "Ar", "x", "y", "z"] and len(vel) == info["n_atoms"] + 1
    rdf = read_table(os.path.join(d, "rdf.dat"))
    assert rdf[0] == ["r", "Ar-Ar"] and len(rdf) == 1 + int(8.0 / 0.1)
    g = np.array([float(x[1]) for x in rdf[1:]])
    assert g[:25].max() == 0.0 and g.max() > 1.5              # excluded core, first shell
    assert "Kin=" in r.stdout and "Vdw=" in r.stdout
    # the same deck through the ABI gives the driver's first row
    lib = abi.load_library()  # the deck parser of the product library is host code
    h = C.c_void_p()
    assert lib.az_deck_load(d.encode(), 7, C.byref(h)) == 0
    e = Engine(lib.az_deck_desc(h).contents, lib=azt.emu_lib(), opts=lib.az_deck_opts(h).contents)
    e.step(0, 1)
    assert abs(e.stats().engTot - etot[0]) <= 1e-5 * abs(etot[0]) + 1e-6
    e.close()
    lib.az_deck_free(h)


def test_driver_full_ewald_deck_reports_the_k_space_energy(tmp_path):
    """'elec pme rc alpha kx ky kz' end to end: C1 (real space) and C2 (reciprocal space) columns are live."""
    d = str(tmp_path)
    decks.nacl_deck(d, 4, rc_vdw=7.0, r_real=7.0, alpha=0.35, elec="pme", jitter=0.4, nstep=6, stat=3, cell_list=7.0)
    r = run_driver(d, "--seed", "5")
    assert r.returncode == 0, r.stdout + r.stderr
    stat = read_table(os.path.join(d, "stat.dat"))
    c1 = [float(x[5]) for x in stat[2:]]
    c2 = [float(x[6]) for x in stat[2:]]
    assert len(c1) == 2 and all(v < -100.0 for v in c1) and all(v > 0.1 for v in c2)
    assert "C1=" in r.stdout and "C2=" in r.stdout


def test_driver_reports_deck_errors_like_the_reference(tmp_path):
    decks.lj_argon_deck(str(tmp_path), 4, nstep=10)
    ctl = os.path.join(str(tmp_path), "control.txt")
    open(ctl, "w").write(open(ctl).read().replace("timestep 0.001 ps\n", ""))
    r = run_driver(str(tmp_path))
    assert r.returncode != 0 and "ERROR[411]" in r.stdout


def test_driver_bonds_angles_deck_and_nuclei_rdf(tmp_path):
    """SURVEY 8f-4 + 8f-1 through the process surface: 'bonds' / 'bond_list' + bonds.txt and 'angles' / 'angle_list' + angles.txt
    add the engBnd / engAngle columns of start_stat (src/cuStat.cu:311-314); 'rdf ... nucl' adds rdf_n.dat in copy_nrdf's
    layout (src/cuStat.cu:696-777).  The deck is written like tests/golden/make_golden.py writes the decks the reference's
    serial build ran, and the first row equals the engine driven through the ABI on the same deck."""
    d = str(tmp_path)
    info = decks.lj_argon_deck(d, 5, nstep=12, stat=4, init_vel="zero", rdf_freq=4, rdf_rmax=8.0, rdf_dr=0.1)
    ctl = open(os.path.join(d, "control.txt")).read().splitlines()
    ctl = [(l + "\tnucl") if l.startswith("rdf") else l for l in ctl]
    open(os.path.join(d, "control.txt"), "w").write("\n".join(ctl) + "\n")
    with open(os.path.join(d, "field.txt"), "a") as f:
        f.write("bonds 2\n1\tAr\tAr\tharm\t5.0\t3.8\tcon\tcon\n2\tAr\tAr\tmors\t0.4\t1.5\t4.0\t0.4\tcon\tcon\nbond_list 1\n")
        f.write("angles 2\n1\tAr\thcos\t6\t-0.3\n2\tAr\thcos\t2.5\t0.45\nangle_list 1\n")
    n = info["n_atoms"]
    blist = [(4 * k, 4 * k + 1, 1 + (k % 2)) for k in range(0, n // 4, 3)] + [(2, 4 * 5 * 5 * 4 + 2, 1)]
    open(os.path.join(d, "bonds.txt"), "w").write("%d\n" % len(blist) + "".join("%d %d %d\n" % b for b in blist))
    alist = [(4 * k, 4 * k + 1, 4 * k + 2, 1 + (k % 2)) for k in range(1, n // 4, 4)]
    open(os.path.join(d, "angles.txt"), "w").write("%d\n" % len(alist) + "".join("%d %d %d %d\n" % a for a in alist))
    r = run_driver(d, "--seed", "7")
    assert r.returncode == 0, r.stdout + r.stderr
    stat = read_table(os.path.join(d, "stat.dat"))
    assert stat[0][7:9] == ["engBnd", "engAngle"] and stat[1][7:9] == ["engBnd, eV", "engAngle, eV"]
    ebnd = [float(x[7]) for x in stat[2:]]
    eang = [float(x[8]) for x in stat[2:]]
    assert len(ebnd) == 3 and all(v > 0.01 for v in ebnd) and all(v > 0.001 for v in eang)
    # rdf_n.dat: one nucleus ("Ar" names itself as its nucleus): the table equals rdf.dat column for column
    rdf = read_table(os.path.join(d, "rdf.dat"))
    nrdf = read_table(os.path.join(d, "rdf_n.dat"))
    assert nrdf[0] == ["r", "Ar-Ar"] and nrdf[1:] == rdf[1:]
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.az_deck_load(d.encode(), 7, C.byref(h)) == 0
    desc = lib.az_deck_desc(h).contents
    assert desc.n_bonds == len(blist) and desc.n_angles == len(alist) and desc.n_bond_types == 3
    e = Engine(desc, lib=azt.emu_lib(), opts=lib.az_deck_opts(h).contents)
    e.step(0, 1)
    st = e.stats()
    assert abs(st.engBond - ebnd[0]) <= 1e-5 * abs(ebnd[0]) and abs(st.engAngle - eang[0]) <= 1e-5 * abs(eang[0])
    o = azt.Oracle(desc)
    o.step(0, 1)
    assert abs(o.stats().engBond_d - st.engBond_d) <= 5e-6 * abs(st.engBond_d)
    e.close()
    lib.az_deck_free(h)
    # mutable bonds are reactive chemistry: refused, not ignored
    with open(os.path.join(d, "field.txt")) as f:
        txt = f.read().replace("3.8\tcon\tcon", "3.8\tcon\tmut\t4.5\t2")
    open(os.path.join(d, "field.txt"), "w").write(txt)
    r = run_driver(d)
    assert r.returncode != 0 and "reactive" in r.stdout


def test_nuclei_rdf_of_a_two_species_deck(tmp_path):
    """Two species, each its own nucleus (second word of the 'spec' lines): rdf_n.dat carries the Na-Na, Na-Cl, Cl-Cl columns of the
    nuclei, normalised by the per-nucleus atom counts (Field::nnumbers), and the deck surface reports the nucleus table."""
    d = str(tmp_path)
    decks.nacl_deck(d, 4, rc_vdw=7.0, r_real=7.0, alpha=0.35, elec="fenn", nstep=4, stat=2, cell_list=7.0)
    ctl = open(os.path.join(d, "control.txt")).read().splitlines()
    ctl = [("rdf\t7.0 0.1\t2\t1000000\tnucl") if l.startswith("rdf") else l for l in ctl]
    open(os.path.join(d, "control.txt"), "w").write("\n".join(ctl) + "\n")
    r = run_driver(d, "--seed", "5")
    assert r.returncode == 0, r.stdout + r.stderr
    nrdf = read_table(os.path.join(d, "rdf_n.dat"))
    rdf = read_table(os.path.join(d, "rdf.dat"))
    assert nrdf[0] == ["r", "Na-Na", "Na-Cl", "Cl-Cl"] and rdf[0] == nrdf[0]
    assert nrdf[1:] == rdf[1:]  # every species is its own nucleus here
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.az_deck_load(d.encode(), 5, C.byref(h)) == 0
    nof = (C.c_int * 2)()
    assert lib.az_deck_nuclei(h, nof) == 2 and list(nof) == [0, 1] and lib.az_deck_rdf_nuclei(h) == 1
    assert lib.az_deck_nucleus_name(h, 1) == b"Cl" and lib.az_deck_nucleus_count(h, 0) == 256
    lib.az_deck_free(h)
