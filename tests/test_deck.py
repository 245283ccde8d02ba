"""Host side of the drop-in: the reference's input files parsed by csrc/az_deck.cpp
(token-scan semantics of src/utils.cpp:87-195, derived values of init_md src/sys_init.cpp:1036-1119)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import azt
from aztotmd_b200 import abi, decks
from aztotmd_b200.engine import KB, M_SCALE, Engine


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()  # loading needs no GPU; only az_engine_create does


def load(lib, d, seed=0):
    h = C.c_void_p()
    rc = lib.az_deck_load(str(d).encode(), seed, C.byref(h))
    return rc, h


CS2_CONTROL = """timestep 0.001 ps
nstep 500000
nequil  10000
eqfreq 2500
temperature 500.0\tradi\t0.2
// nose 0.2  (a comment: unknown words are skipped)
init_vel\tzero\t0.0332

permittivity  1.0
cell_list\t2.7
max_neigh\t185

elec\tfenn\t8.0\t0.4\t6\t6\t6

rdf\t8.0   0.02\t10\t5000\t-nucl
eJump\t0\t1.7\tmetr
Ux\t\t0.0
revcon\t\t50000
stat\t\t200
"""
CS2_FIELD = """spec 1
Ar\tAr\t39.9\t0.0\t0.0
red-ox 0
vdw 1
Ar\tAr\tsurk\t6.0\t75.0\t8.0\t1.0\t1.0
radii 1
Ar\t2.73\t4.731\t0.2


free text after the sections, like the shipped deck: the first 'vdw'/'radii' match wins
vdw 1
Ar\tAr\tlnjs\t4.0    0.01006\t3.3952
"""


def write_cs2_like(tmp, crlf=False, n_side=5):
    g = np.arange(n_side) * 3.0 + 10.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    decks.write_deck(str(tmp), names="Ar", pos=pos, box=(35.0, 35.0, 35.0), field_txt=CS2_FIELD, control_txt=CS2_CONTROL,
                     cuda_txt="nstep stat 50\nnstep msdstat 50\nnthread a 16\nnthread b 32\n")
    if crlf:
        for f in ("control.txt", "field.txt", "cuda.txt", "atoms.xyz"):
            p = os.path.join(str(tmp), f)
            data = open(p).read().replace("\n", "\r\n")
            open(p, "w", newline="").write(data)
    return len(pos)


def test_lj_deck_fields(lib, tmp_path):
    info = decks.lj_argon_deck(str(tmp_path), 5, nstep=77, stat=11, rebuild_every=10, skin=0.4)
    rc, h = load(lib, tmp_path, seed=5)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    o = lib.az_deck_opts(h).contents
    assert d.n_atoms == 500 and d.n_species == 1 and lib.az_deck_nsteps(h) == 77 and lib.az_deck_stat_period(h) == 11
    assert (d.la, d.lb, d.lc) == pytest.approx(tuple(np.round(info["box"], 6)))
    assert d.species[0].mass == pytest.approx(39.9 * M_SCALE, rel=1e-12) and d.species[0].name == b"Ar"
    v = d.vdw[0]
    assert v.kind == abi.VDW_LJ and v.p0 == pytest.approx(4 * 0.01006) and v.p1 == pytest.approx(3.3952 ** 2) and v.p2 == pytest.approx(24 * 0.01006)
    assert v.r2cut == pytest.approx(8.488 ** 2) and d.r2_max == pytest.approx(8.488 ** 2) and d.cell_size == pytest.approx(8.488)
    assert d.elec_kind == abi.ELEC_NONE and d.tstat_kind == abi.TSTAT_NONE
    assert d.t_kin == pytest.approx(0.5 * 86.6 * KB * 1500)  # no thermostat: degFree = 3N
    assert o.rebuild_every == 10 and o.skin == pytest.approx(0.4)
    # 'init_vel gaus': zero total momentum, kinetic energy scaled to tKin (gauss_temp)
    n = d.n_atoms
    vel = np.stack([np.ctypeslib.as_array(p, (n,)) for p in (d.vxs, d.vys, d.vzs)], 1)
    assert np.abs(vel.sum(0)).max() < 1e-9 * np.abs(vel).sum()
    assert 0.5 * d.species[0].mass * (vel ** 2).sum() == pytest.approx(d.t_kin, rel=1e-12)
    lib.az_deck_free(h)


def test_seed_makes_velocities_reproducible(lib, tmp_path):
    decks.lj_argon_deck(str(tmp_path), 4, nstep=1)
    out = []
    for seed in (3, 3, 4):
        rc, h = load(lib, tmp_path, seed=seed)
        d = lib.az_deck_desc(h).contents
        out.append(np.ctypeslib.as_array(d.vxs, (d.n_atoms,)).copy())
        lib.az_deck_free(h)
    assert np.array_equal(out[0], out[1]) and not np.array_equal(out[0], out[2])


@pytest.mark.parametrize("crlf", [False, True], ids=["lf", "crlf"])
def test_case_study_2_like_deck(lib, tmp_path, crlf):
    n = write_cs2_like(tmp_path, crlf=crlf)
    rc, h = load(lib, tmp_path, seed=9)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    assert d.n_atoms == n and lib.az_deck_nsteps(h) == 500000 and lib.az_deck_stat_period(h) == 200
    assert d.n_equil == 10000 and d.eq_freq == 2500
    assert d.tstat_kind == abi.TSTAT_RADI and d.temperature == 500.0
    assert d.t_kin == pytest.approx(0.5 * 500.0 * KB * (3 * n - 1))   # thermostat on: one degree of freedom less
    # uncharged species: 'elec fenn' is downgraded to none, the vdW cutoff rules (src/elec.cpp:52-57, sys_init.cpp:1059-1072)
    assert d.elec_kind == abi.ELEC_NONE and d.r_real == 8.0 and d.r2_max == pytest.approx(36.0)
    v = d.vdw[0]  # first 'vdw' section wins
    assert v.kind == abi.VDW_SURK and v.use_radii == 1 and (v.p0, v.p1, v.p2, v.p3) == (75.0, 8.0, 1.0, 1.0)
    assert (d.species[0].radA, d.species[0].radB, d.species[0].mxEng) == (2.73, 4.731, 0.2)
    assert d.cell_size == pytest.approx(2.7)
    ph = np.ctypeslib.as_array(d.photons, (n,))
    assert (ph > 0).all() and (ph < 1).all()
    assert b"WARNING[b004]" in lib.az_deck_warnings(h)
    lib.az_deck_free(h)


def test_photon_energies_follow_e4_exp_distribution(lib, tmp_path):
    """photon_engs inverts the CDF of P(e) ~ e^4 exp(-e/kT) (mean 5 kT) by bisection with an absolute
    tolerance on the residual and at most 20 halvings; energetic draws that do not converge reuse the
    previous sample (src/temperature.cpp:70-80).  The restatement keeps that, so the sample mean sits a
    few per cent BELOW 5 kT -- exactly like the reference's own table."""
    write_cs2_like(tmp_path, n_side=14)
    rc, h = load(lib, tmp_path, seed=1)
    d = lib.az_deck_desc(h).contents
    ph = np.ctypeslib.as_array(d.photons, (d.n_atoms,))
    kt = KB * 500.0
    assert 0.85 * 5 * kt < ph.mean() < 1.0 * 5 * kt
    assert ph.std() == pytest.approx(math.sqrt(5) * kt, rel=0.25)
    lib.az_deck_free(h)


def test_nose_hoover_and_reset_vels_directives(lib, tmp_path):
    write_cs2_like(tmp_path)
    p = os.path.join(str(tmp_path), "control.txt")
    text = open(p).read().replace("radi\t0.2", "nose 0.35", 1) + "\nreset_vels 40\n"
    open(p, "w").write(text)
    rc, h = load(lib, tmp_path)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    assert d.tstat_kind == abi.TSTAT_NOSE and d.tstat_tau == pytest.approx(0.35) and d.reset_vels == 40
    assert not d.photons  # no photon table without the radiative thermostat
    lib.az_deck_free(h)


@pytest.mark.parametrize("edit,status,needle", [
    (("control.txt", "timestep 0.001 ps", ""), abi.AZ_ERR_IO, b"ERROR[411]"),
    (("control.txt", "radi\t0.2", "nose zzz"), abi.AZ_ERR_IO, b"tau"),
    (("control.txt", "elec\tfenn", "elec\tzzz"), abi.AZ_ERR_IO, b"ERROR[404]"),
    (("control.txt", "init_vel\tzero", "no_vel zero"), abi.AZ_ERR_IO, b"ERROR[406]"),
    (("control.txt", "cell_list\t2.7", ""), abi.AZ_ERR_IO, b"cell_list"),
    (("field.txt", "red-ox 0", "bonds 1\n1 Ar Ar harm 5.0 3.8 con mut 4.5 2"), abi.AZ_ERR_UNSUPPORTED, b"reactive"),  # mutable bonds: reactive chemistry
    (("field.txt", "red-ox 0", "linkage 2"), abi.AZ_ERR_UNSUPPORTED, b"linkage"),
    (("field.txt", "surk\t6.0\t75.0\t8.0\t1.0\t1.0", "quux\t6.0\t75.0\t8.0"), abi.AZ_ERR_IO, b"ERROR[006]"),
    (("atoms.xyz", "Ar\t10.000000\t10.000000\t10.000000", "Xe\t10.000000\t10.000000\t10.000000"), abi.AZ_ERR_IO, b"ERROR[009]"),
])
def test_deck_errors_carry_the_reference_codes(lib, tmp_path, edit, status, needle):
    write_cs2_like(tmp_path)
    fname, old, new = edit
    p = os.path.join(str(tmp_path), fname)
    text = open(p).read()
    assert old in text
    open(p, "w").write(text.replace(old, new, 1))
    rc, h = load(lib, tmp_path)
    assert rc == status
    assert needle in lib.az_deck_error(h)
    lib.az_deck_free(h)


def test_missing_file(lib, tmp_path):
    rc, h = load(lib, tmp_path)
    assert rc == abi.AZ_ERR_IO and b"ERROR[001]" in lib.az_deck_error(h)
    lib.az_deck_free(h)


def test_ionic_deck_dsf_constants(lib, tmp_path):
    decks.nacl_deck(str(tmp_path), 3, r_real=12.0, alpha=0.2, elec="fenn")
    rc, h = load(lib, tmp_path)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    assert d.elec_kind == abi.ELEC_FENNEL and d.r2_max == pytest.approx(144.0)  # with electrostatics rMax = rReal only
    arc = 0.2 * 12.0
    daipi2 = 2 * 0.2 / math.sqrt(3.14159265359)
    assert d.daipi2 == pytest.approx(daipi2, rel=1e-12)
    assert d.el_c1 == pytest.approx(math.erfc(arc) / 12.0, rel=1e-12)
    assert d.el_c2 == pytest.approx(math.erfc(arc) / 144.0 + daipi2 * math.exp(-arc * arc) / 12.0, rel=1e-12)
    ns = d.n_species
    assert d.vdw[0 * ns + 1].kind == abi.VDW_BUCK and d.vdw[1 * ns + 0].p0 == d.vdw[0 * ns + 1].p0 == 1256.31
    lib.az_deck_free(h)


def test_ionic_deck_pme_carries_the_k_ranges(lib, tmp_path):
    """'elec pme rc alpha kx ky kz' (src/elec.cpp:47-51): the k ranges reach the engine, which derives the k-space
    constants itself; the deck-built description steps through the kernels (emulator) like the oracle."""
    import parity_cases
    decks.nacl_deck(str(tmp_path), 5, rc_vdw=9.0, r_real=9.0, alpha=0.3, elec="pme", jitter=0.4, nstep=2)
    rc, h = load(lib, tmp_path, seed=3)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    assert d.elec_kind == abi.ELEC_EWALD and list(d.ewald_k) == [6, 6, 6] and d.alpha == 0.3
    assert d.daipi2 == pytest.approx(2 * 0.3 / math.sqrt(3.14159265359), rel=1e-12)
    o = azt.Oracle(d)
    e = Engine(d, lib=azt.emu_lib())
    e.set_option("energy_every_step", 1)
    o.step(0, 1); e.step(0, 1)
    do, de = o.dump(), e.dump()
    parity_cases.assert_cells_equal(do, de)
    parity_cases.assert_forces_close(o, do, de)
    parity_cases.assert_stats_close(o.stats(), e.stats())
    assert e.stats().engCoul2 > 1.0
    e.close()
    lib.az_deck_free(h)


def test_deck_drives_the_kernels_like_the_python_system(lib, tmp_path):
    """A deck-built description goes through the same kernels (emulator build) and matches the oracle."""
    import parity_cases
    decks.lj_argon_deck(str(tmp_path), 5, nstep=3)
    rc, h = load(lib, tmp_path, seed=7)
    assert rc == 0
    d = lib.az_deck_desc(h).contents
    o = azt.Oracle(d)
    e = Engine(d, lib=azt.emu_lib())
    for step in range(2):
        o.step(step, 1); e.step(step, 1)
        do, de = o.dump(), e.dump()
        parity_cases.assert_cells_equal(do, de)
        parity_cases.assert_forces_close(o, do, de)
    e.close()
    lib.az_deck_free(h)


DECKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decks")


@pytest.mark.parametrize("case,n,cells", [("cs1", 40000, 13), ("cs2", 4000, 12)])
def test_shipped_case_studies_parse(lib, case, n, cells):
    """tests/golden/decks/cs1, cs2: the reference's two shipped decks (input fixtures, byte copies)."""
    rc, h = load(lib, os.path.join(DECKS, case), seed=1)
    assert rc == 0, lib.az_deck_error(h)
    d = lib.az_deck_desc(h).contents
    assert d.n_atoms == n and d.tstat_kind == abi.TSTAT_RADI
    assert int(math.floor(np.float32(d.la) / np.float32(d.cell_size))) == cells
    lib.az_deck_free(h)


@pytest.mark.skipif(not os.path.isdir("/root/reference/case study 1"), reason="reference tree not present")
@pytest.mark.parametrize("case,ref", [("cs1", "case study 1"), ("cs2", "case study 2")])
def test_committed_decks_are_the_shipped_files(case, ref):
    for f in ("atoms.xyz", "control.txt", "field.txt", "cuda.txt"):
        assert open(os.path.join(DECKS, case, f), "rb").read() == open(os.path.join("/root/reference", ref, f), "rb").read(), f


def test_oracle_reproduces_the_reference_serial_build_on_case_study_1(lib, tmp_path):
    """Pins oracle/az_oracle.c on the shipped 40 000-atom deck: same velocities as the reference's serial build drew
    (tests/golden/deck_cs1_ref_serial.npz), 20 steps, its kinetic energy, the one pair that comes inside rc, positions."""
    import shutil
    g = np.load(os.path.join(os.path.dirname(DECKS), "deck_cs1_ref_serial.npz"))
    d = os.path.join(str(tmp_path), "cs1")
    shutil.copytree(os.path.join(DECKS, "cs1"), d)
    ctl = open(os.path.join(d, "control.txt"), encoding="latin-1").read().replace("temperature 298.0\tradi\t0.2", "temperature 298.0\tnone")
    open(os.path.join(d, "control.txt"), "w", encoding="latin-1").write(ctl)
    rc, h = load(lib, d, seed=1)
    assert rc == 0
    desc = lib.az_deck_desc(h).contents
    n, L = desc.n_atoms, float(desc.la)
    pos = np.stack([np.ctypeslib.as_array(p, shape=(n,)) for p in (desc.xs, desc.ys, desc.zs)], 1).copy()
    pos0 = (pos + g["shift"]) % L
    keep = [np.ascontiguousarray(pos0[:, a]) for a in range(3)] + [np.ascontiguousarray(g["vel0"][:, a]) for a in range(3)]
    ptr = [a.ctypes.data_as(C.POINTER(C.c_double)) for a in keep]
    desc.xs, desc.ys, desc.zs, desc.vxs, desc.vys, desc.vzs = ptr
    o = azt.Oracle(desc)
    o.step(0, int(g["nstep"]))
    st = o.stats()
    assert abs(st.engKin_d - float(g["engKin"])) <= 2e-6 * float(g["engKin"])
    a = o.download(abi.ORDER_ORIGINAL)
    dp = a["pos"][::16] - g["pos_sub"]
    dp -= L * np.round(dp / L)
    assert np.abs(dp).max() <= 1.3e-3   # 20 steps x half an ulp of an fp32 coordinate in [1024, 1141.5) A (ulp = 1.2e-4 A): the reference's CUDA build holds positions in fp32 too
    dmp = o.dump()
    f = np.zeros((n, 3)); f[dmp["ids"]] = dmp["frc"]
    nz = np.nonzero(np.abs(f).max(1) > 0)[0]
    assert np.array_equal(nz, g["frc_idx"])
    assert np.abs(f[nz] - g["frc"]).max() <= 2e-3 * np.abs(g["frc"]).max()
    lib.az_deck_free(h)
