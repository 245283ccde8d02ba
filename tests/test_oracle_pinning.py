"""Pin the CPU oracle (oracle/az_oracle.c) before anything trusts it.

The reference ships no golden vectors, so the oracle is pinned against
  (a) forces/energies produced by the reference's OWN serial build (oracle/_ref, compiled from
      /root/reference/src) and committed as tests/golden/ref_serial_*.npz (tests/golden/make_golden.py);
  (b) the analytic known answers listed in SURVEY.md 8(c);
  (c) an independent O(N^2) minimum-image evaluation of the same pair functions.
"""
import ast
import ctypes as C
import glob
import math
import os

import numpy as np
import pytest

import azt
from aztotmd_b200 import abi
from aztotmd_b200.engine import KB, M_SCALE, Species, System

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_fp = C.POINTER(C.c_float)


def system_from_fixture(path):
    z = np.load(path)
    meta = ast.literal_eval(str(z["meta"]))
    pos, types = z["pos"], z["types"]
    if meta["kind"] == "lj":
        s = System(pos=pos, types=types, box=meta["box"], species=[Species("Ar", azt.AR_MASS)], cell_size=meta["rc"])
        s.set_vdw(0, 0, "lnjs", meta["rc"], [azt.AR_EPS, azt.AR_SIGMA])
    else:
        ek = {"fenn": abi.ELEC_FENNEL, "dir": abi.ELEC_DIRECT, "pme": abi.ELEC_EWALD}[meta["elec"]]
        s = System(pos=pos, types=types, box=meta["box"], species=[Species("Na", 22.99, 1.0), Species("Cl", 35.45, -1.0)],
                   cell_size=max(meta["r_real"], meta["rc_vdw"]), elec_kind=ek, r_real=meta["r_real"], alpha=meta["alpha"],
                   ewald_k=meta.get("ewald_k", (0, 0, 0)))
        s.set_vdw(0, 0, "buck", meta["rc_vdw"], [424.097, 0.317, 1.05])
        s.set_vdw(0, 1, "buck", meta["rc_vdw"], [1256.31, 0.317, 7.00])
        s.set_vdw(1, 1, "buck", meta["rc_vdw"], [3488.998, 0.317, 72.40])
    if "bonds" in meta:  # the fixture's constant bonds / valence angles travel in the system description
        names = {v: k for k, v in abi.BOND_BY_NAME.items()}
        s.bond_types = [(names[kind], list(p)) for kind, p in meta["bond_types"]]
        s.bonds = np.array(meta["bonds"], dtype=np.int32)
        if "angles" in meta:
            s.angle_types = [tuple(t) for t in meta["angle_types"]]
            s.angles = np.array(meta["angles"], dtype=np.int32)
    return s, z


def bonded_forces(meta, pos):
    """oracle/az_oracle.c: azo_apply_const_bonds on the fixture's bond list; returns (forces[n, 3], bond energy)"""
    lib = azt.oracle_lib()
    ip, fp, dp = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.azo_apply_const_bonds.restype = C.c_double
    lib.azo_apply_const_bonds.argtypes = [C.c_int, ip, ip, ip, ip, fp, C.c_int, fp, fp, fp, fp, dp, dp, dp]
    b = np.array(meta["bonds"], dtype=np.int32)
    at1, at2, bt = (np.ascontiguousarray(b[:, k]) for k in range(3))
    nt = max(t[0] for t in meta["bond_types"]) + 1
    kinds = np.zeros(nt, dtype=np.int32)
    par = np.zeros((nt, 5), dtype=np.float32)
    for k, (kind, p) in enumerate(meta["bond_types"]):  # type ids of the list start at 1 ([0] = empty bond)
        kinds[k + 1] = kind
        par[k + 1] = p
    n = len(pos)
    xyz = [np.ascontiguousarray(pos[:, a], dtype=np.float32) for a in range(3)]
    L = np.array(meta["box"], dtype=np.float32)
    f = [np.zeros(n) for _ in range(3)]
    eng = lib.azo_apply_const_bonds(len(b), at1.ctypes.data_as(ip), at2.ctypes.data_as(ip), bt.ctypes.data_as(ip), kinds.ctypes.data_as(ip),
                                    par.ctypes.data_as(fp), n, *(a.ctypes.data_as(fp) for a in xyz), L.ctypes.data_as(fp),
                                    *(a.ctypes.data_as(dp) for a in f))
    return np.stack(f, 1), eng


def angle_forces(meta, pos):
    """oracle/az_oracle.c: azo_apply_angles on the fixture's angle list; returns (forces[n, 3], angle energy)"""
    lib = azt.oracle_lib()
    ip, fp, dp = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.azo_apply_angles.restype = C.c_double
    lib.azo_apply_angles.argtypes = [C.c_int, ip, ip, ip, ip, fp, fp, fp, fp, fp, dp, dp, dp]
    a = np.array(meta["angles"], dtype=np.int32)
    cols = [np.ascontiguousarray(a[:, k]) for k in range(4)]
    par = np.zeros((len(meta["angle_types"]) + 1, 2), dtype=np.float32)
    for k, t in enumerate(meta["angle_types"]):
        par[k + 1] = t
    n = len(pos)
    xyz = [np.ascontiguousarray(pos[:, k], dtype=np.float32) for k in range(3)]
    L = np.array(meta["box"], dtype=np.float32)
    f = [np.zeros(n) for _ in range(3)]
    eng = lib.azo_apply_angles(len(a), *(c.ctypes.data_as(ip) for c in cols), par.ctypes.data_as(fp), *(v.ctypes.data_as(fp) for v in xyz),
                               L.ctypes.data_as(fp), *(v.ctypes.data_as(dp) for v in f))
    return np.stack(f, 1), eng


FIXTURES = sorted(glob.glob(os.path.join(GOLD, "ref_serial_*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[11:-4] for p in FIXTURES])
def test_oracle_matches_reference_serial_build(path):
    """float32 oracle vs the reference's double-precision serial forces: 1e-5 relative to the
    per-atom sum of |pair terms| (the fp32 noise floor), energies to 2e-6 relative."""
    s, z = system_from_fixture(path)
    o = azt.Oracle(s)
    o.forces_only()
    d = azt.sort_by_id(o.dump())
    scale = np.zeros(o.n, dtype=np.float32)
    scale[o.dump()["ids"]] = o.force_scale()
    frc = d["frc"].astype(np.float64)
    meta = ast.literal_eval(str(z["meta"]))
    st = o.stats()
    if "bonds" in meta:  # constant bonds on top of the pair forces (apply_const_bonds / exec_bondlist), inside the oracle's step
        fb, eb = bonded_forces(meta, z["pos"])  # ... and the executor on its own
        assert np.abs(fb).max(1).astype(bool).sum() == len({a for b in meta["bonds"] for a in b[:2]})
        assert eb == pytest.approx(float(z["engBond"]), rel=5e-6)
        assert st.engBond_d == pytest.approx(float(z["engBond"]), rel=5e-6)
        if "angles" in meta:  # valence angles (apply_angles / exec_anglelist)
            fa, ea = angle_forces(meta, z["pos"])
            assert abs(ea) > 1e-3 and ea == pytest.approx(float(z["engAngle"]), rel=5e-6)
            assert st.engAngle_d == pytest.approx(float(z["engAngle"]), rel=5e-6)
    err = np.abs(frc - z["frc"]).max(1)
    assert (err <= 1e-5 * np.maximum(scale, 1e-12)).all(), float((err / np.maximum(scale, 1e-12)).max())
    assert st.engVdW_d == pytest.approx(float(z["engVdW"]), rel=2e-6)
    # the serial twin's direct/DSF real-space energy (Sim::engElec3)
    if float(z["engCoul"]) != 0.0:
        assert st.engCoul1_d == pytest.approx(float(z["engCoul"]), rel=5e-6)
    if "engElec2" in z.files:  # 'elec pme': reciprocal-space energy (Sim::engElec2) and the constant term (Sim::engElec1)
        assert st.engCoul2_d == pytest.approx(float(z["engElec2"]), rel=5e-6)
        assert st.engCoul3 == pytest.approx(float(z["engElec1"]), rel=5e-6)


def pair(kind, p, r, rc=100.0, elec=0, alpha=0.0, q=0.0, rad=(0.0, 0.0), r_real=0.0):
    lib = azt.oracle_lib()
    pp = (C.c_float * 5)(*(list(p) + [0.0] * (5 - len(p))))
    ev, ec = C.c_float(), C.c_float()
    daipi2 = 2 * alpha / math.sqrt(3.14159265359)
    c1 = c2 = 0.0
    if elec == abi.ELEC_FENNEL:
        arc = alpha * r_real
        c1 = math.erfc(arc) / r_real
        c2 = math.erfc(arc) / r_real ** 2 + daipi2 * math.exp(-arc * arc) / r_real
    rmax2 = max(rc, r_real) ** 2 if elec else rc * rc
    f = lib.azo_pair_eval(kind, pp, rc * rc, elec, alpha, daipi2, c1, c2, r_real, q, r * r, rmax2, rad[0], rad[1],
                          C.byref(ev), C.byref(ec))
    return f, ev.value, ec.value


def test_lj_known_answers():
    eps, sig = azt.AR_EPS, azt.AR_SIGMA
    p = [4 * eps, sig * sig, 24 * eps]
    f, e, _ = pair(abi.VDW_LJ, p, sig)                       # U(sigma) = 0, F/r = 24 eps / sigma^2
    assert abs(e) < 1e-8 and f == pytest.approx(24 * eps / sig ** 2, rel=1e-5)
    f, e, _ = pair(abi.VDW_LJ, p, 2 ** (1 / 6) * sig)        # minimum: U = -eps, F = 0
    assert e == pytest.approx(-eps, rel=1e-5) and abs(f) < 1e-7
    f, e, _ = pair(abi.VDW_LJ, p, 2.5 * sig, rc=2.5 * sig + 1e-4)  # at the cutoff: plain truncation
    assert e == pytest.approx(4 * eps * (2.5 ** -12 - 2.5 ** -6), rel=1e-4)
    f, e, _ = pair(abi.VDW_LJ, p, 2.5 * sig + 0.01, rc=2.5 * sig)  # beyond it: nothing
    assert f == 0.0 and e == 0.0


def test_autocap_clamps_f_over_r():
    eps, sig = azt.AR_EPS, azt.AR_SIGMA
    f, _, _ = pair(abi.VDW_LJ, [4 * eps, sig * sig, 24 * eps], 0.5)  # deep overlap -> MX_FRC (src/cuPairs.cu:94-115)
    assert f == abi.AZ_MX_FRC


def test_dsf_force_vanishes_at_cutoff():
    f, _, ec = pair(abi.VDW_NONE, [0] * 5, 12.0 - 1e-4, elec=abi.ELEC_FENNEL, alpha=0.2, q=-1.0, r_real=12.0)
    assert abs(f) < 1e-6 and abs(ec) < 1e-6
    f2, _, _ = pair(abi.VDW_NONE, [0] * 5, 3.0, elec=abi.ELEC_FENNEL, alpha=0.2, q=-1.0, r_real=12.0)
    assert f2 < 0  # opposite charges attract (F/r negative)


def test_direct_coulomb_known_answer():
    f, _, ec = pair(abi.VDW_NONE, [0] * 5, 2.0, elec=abi.ELEC_DIRECT, q=1.0, r_real=9.0)
    assert ec == pytest.approx(14.3996 / 2.0, rel=1e-6) and f == pytest.approx(14.3996 / 8.0, rel=1e-6)


def test_buckingham_with_and_without_coulomb_agree():
    p = [1256.31, 0.317, 7.0]
    f0, e0, _ = pair(abi.VDW_BUCK, p, 2.8, rc=9.0)
    f1, e1, ec = pair(abi.VDW_BUCK, p, 2.8, rc=9.0, elec=abi.ELEC_DIRECT, q=0.0, r_real=9.0)
    assert f0 == pytest.approx(f1, rel=1e-6) and e0 == pytest.approx(e1, rel=1e-6) and ec == 0.0
    r = 2.8
    assert e0 == pytest.approx(p[0] * math.exp(-r / p[1]) - p[2] / r ** 6, rel=1e-5)
    assert f0 == pytest.approx(p[0] * math.exp(-r / p[1]) / r / p[1] - 6 * p[2] / r ** 8, rel=1e-5)


def test_bmh_elin_einv_known_answers():
    """The three remaining analytic forms of the table (src/cuVdW.cu:115-136,162-179,200-217) against their formulas."""
    r = 2.9
    A, B, sg, Cc, D = 0.25, 3.0, 3.2, 70.0, 140.0  # bmhs: U = A exp[B(s - r)] - C/r^6 - D/r^8
    f, e, _ = pair(abi.VDW_BMH, [A, B, sg, Cc, D], r, rc=9.0)
    assert e == pytest.approx(A * math.exp(B * (sg - r)) - Cc / r ** 6 - D / r ** 8, rel=2e-5)
    assert f == pytest.approx(A * B * math.exp(B * (sg - r)) / r - 6 * Cc / r ** 8 - 8 * D / r ** 10, rel=2e-5)
    A, ro, Cc = 900.0, 0.33, 0.02                   # elin: U = A exp(-r/ro) + C r
    f, e, _ = pair(abi.VDW_ELIN, [A, ro, Cc], r, rc=9.0)
    assert e == pytest.approx(A * math.exp(-r / ro) + Cc * r, rel=2e-5)
    assert f == pytest.approx(A * math.exp(-r / ro) / r / ro - Cc / r, rel=2e-5)
    A, ro, Cc = 900.0, 0.33, 1.5                    # einv: U = A exp(-r/ro) - C/r
    f, e, _ = pair(abi.VDW_EINV, [A, ro, Cc], r, rc=9.0)
    assert e == pytest.approx(A * math.exp(-r / ro) - Cc / r, rel=2e-5)
    assert f == pytest.approx(A * math.exp(-r / ro) / r / ro - Cc / r ** 3, rel=2e-5)


def test_surk_at_initial_radii():
    # case study 2: surk 6.0 75.0 8.0 1.0 1.0 with a = b = 2.73/4.731 (radii at U = 0)
    a = 2.73 / 4.731
    assert a == pytest.approx(0.5770, abs=1e-4)
    c1, c2, k, l = 75.0, 8.0, 1.0, 1.0
    r = 3.0
    f, e, _ = pair(abi.VDW_SURK, [c1, c2, k, l], r, rc=6.0, rad=(a, a))
    assert e == pytest.approx(a * a / r ** 6 * (c1 * a ** 4 / r - c2 / (k * a + l * a)), rel=1e-5)
    assert f == pytest.approx(a * a / r ** 8 * (7 * c1 * a ** 4 / r - 6 * c2 / (k * a + l * a)), rel=1e-5)


def test_grid_of_the_case_studies():
    lib = azt.oracle_lib()
    cn = (C.c_int * 3)(); cs = (C.c_float * 3)(); cr = (C.c_float * 3)()
    lib.azo_split_cells(1141.5, 1141.5, 1141.5, 85.0, cn, cs, cr)  # case study 1: 13^3 cells of 87.81
    assert tuple(cn) == (13, 13, 13) and cs[0] == pytest.approx(87.8077, rel=1e-6)
    lib.azo_split_cells(35.0, 35.0, 35.0, 2.7, cn, cs, cr)         # case study 2: 12^3 cells
    assert tuple(cn) == (12, 12, 12)
    assert cr[0] == np.float32(12) / np.float32(35.0)


def test_cell_index_uses_double_floor_and_clamps():
    lib = azt.oracle_lib()
    cn = (C.c_int * 3)(3, 3, 3); cr = (C.c_float * 3)(*[np.float32(3) / np.float32(28.5)] * 3)
    assert lib.azo_cell_index(0.0, 0.0, 0.0, cr, cn) == 0
    assert lib.azo_cell_index(28.4999, 0.0, 9.6, cr, cn) == 2 * 9 + 1
    assert lib.azo_cell_index(28.5, 28.5, 28.5, cr, cn) == 26  # x*crev rounds to n: clamped (reference has no clamp)


def test_tkin_formula():
    s = azt.lj_system(ncell=(3, 3, 3), temperature=298.0, tstat=abi.TSTAT_RADI, vel=False)
    n = len(s.pos)
    assert s.t_kin == pytest.approx(0.5 * 298.0 * KB * (3 * n - 1))
    assert M_SCALE == pytest.approx(1.0364e-4, rel=1e-3)


def test_unit_vectors_are_unit_and_paired():
    lib = azt.oracle_lib()
    ux, uy, uz = (np.zeros(abi.AZ_N_UVECT) for _ in range(3))
    dp = C.POINTER(C.c_double)
    lib.azo_default_uvects(ux.ctypes.data_as(dp), uy.ctypes.data_as(dp), uz.ctypes.data_as(dp))
    nrm = np.sqrt(ux ** 2 + uy ** 2 + uz ** 2)
    assert np.allclose(nrm, 1.0, atol=1e-9)
    assert np.allclose(ux[0::2], -ux[1::2]) and np.allclose(uz[0::2], -uz[1::2])


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10: counter = key = 0
    lib = azt.oracle_lib()
    out = (C.c_uint32 * 4)()
    lib.azo_philox(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    lib.azo_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


@pytest.mark.parametrize("maker", [lambda: azt.lj_system(ncell=(5, 5, 5)), lambda: azt.nacl_system(ncell=(5, 5, 5)),
                                   lambda: azt.nacl_system(ncell=(5, 5, 5), kind="mixed", elec=abi.ELEC_EWALD),
                                   lambda: azt.surk_system(n_side=6, tstat=abi.TSTAT_NONE)],
                         ids=["lj", "buck_dsf", "mixed_ewald", "surk"])
def test_cell_traversal_equals_all_pairs(maker):
    """The stencil traversal (reference image rule) against a plain minimum-image O(N^2) sweep."""
    s = maker()
    o = azt.Oracle(s)
    o.forces_only()
    d = o.dump()
    f2, ev, ec = o.forces_allpairs()
    sc = np.maximum(o.force_scale(), 1e-12)
    assert (np.abs(d["frc"] - f2).max(1) <= 2e-6 * sc).all()
    st = o.stats()
    assert st.engVdW_d == pytest.approx(ev, rel=1e-9, abs=1e-9) and st.engCoul1_d == pytest.approx(ec, rel=1e-9, abs=1e-9)


def test_case_study_1_has_no_pair_inside_cutoff():
    """in the shipped 40 000-atom deck the minimum distance (4.008 A) exceeds rc = 4.0 A,
    so the initial potential energy is exactly zero.  Checked on a 4 000-atom cut of the same recipe:
    a dilute gas placed on a jittered lattice with spacing > rc."""
    rng = np.random.default_rng(7)
    g = (np.arange(16) + 0.5) * 6.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:4000] + rng.uniform(-0.9, 0.9, (4000, 3))
    s = System(pos=pos, types=np.zeros(4000, dtype=np.int32), box=(96.0, 96.0, 96.0), species=[Species("Ar", 39.9)],
               cell_size=8.0)
    s.set_vdw(0, 0, "lnjs", 4.0, [0.01006, 3.3952])
    o = azt.Oracle(s)
    o.forces_only()
    assert o.pairs_in_range() == 0 and o.stats().engVdW_d == 0.0 and not o.dump()["frc"].any()


# ---------------------------------------------------------------------------------------------------------------
# The reference's own __device__ functions, executed on the CPU (oracle/ref_device_harness.cpp: bodies cut verbatim from
# src/cuVdW.cu, src/cuElec.cu, src/cuTemp.cu at build time) -> tests/golden/ref_device_vectors.npz
# (tests/golden/make_golden_device.py).  These pin the arithmetic that exists only in the CUDA build: surk and the
# radiative thermostat, plus every other pair function including the *_r variants.
DEV = os.path.join(GOLD, "ref_device_vectors.npz")


def test_oracle_pair_functions_match_reference_device_source():
    z = np.load(DEV)
    lib = azt.oracle_lib()
    worst = 0.0
    for row in z["vdw"]:
        kind, with_r, r2, ra, rb, f_ref, e_ref = int(row[0]), int(row[1]), row[2], row[3], row[4], row[5], row[6]
        p = (C.c_float * 5)(*row[7:12])
        ev, ec = C.c_float(), C.c_float()
        # with_r: the reference takes r from the Coulomb term that ran first (pair_1, src/cuPairs.cu:131-146); a direct
        # Coulomb term with zero charge product supplies r without adding anything
        elec = abi.ELEC_DIRECT if with_r else abi.ELEC_NONE
        f = lib.azo_pair_eval(kind, p, 1e6, elec, 0.0, 0.0, 0.0, 0.0, 1e3, 0.0, r2, 1e6, ra, rb, C.byref(ev), C.byref(ec))
        if abs(f_ref) <= abi.AZ_MX_FRC:  # force_autocap clamps the total, the device function does not
            worst = max(worst, abs(f - f_ref) / max(abs(f_ref), 1e-12))
            assert f == pytest.approx(f_ref, rel=3e-6, abs=1e-12), (kind, with_r, r2)
        assert ev.value == pytest.approx(e_ref, rel=3e-6, abs=1e-10), (kind, with_r, r2)
    alpha, daipi2, c1, c2, rc = z["coul_consts"]
    for kind, r2, q, f_ref, e_ref, r_ref in z["coul"]:
        ev, ec = C.c_float(), C.c_float()
        p = (C.c_float * 5)()
        f = lib.azo_pair_eval(abi.VDW_NONE, p, 0.0, int(kind), alpha, daipi2, c1, c2, rc, q, r2, 1e6, 0.0, 0.0, C.byref(ev), C.byref(ec))
        assert f == pytest.approx(f_ref, rel=3e-6, abs=1e-9), (kind, r2)
        assert ec.value == pytest.approx(e_ref, rel=3e-6, abs=2e-7), (kind, r2)
    assert worst < 3e-6


def test_oracle_angled_vector_matches_reference_device_source():
    z = np.load(DEV)
    lib = azt.oracle_lib()
    for row in z["angled"]:
        v = (C.c_float * 3)(*row[0:3])
        o = (C.c_float * 3)()
        lib.azo_angled_vector(v, C.c_float(row[3]), C.c_float(row[4]), o)
        assert np.allclose(np.array(o[:]), row[5:8], rtol=2e-5, atol=2e-6), row


def test_oracle_thermostat_atom_matches_reference_device_source():
    """adsorb_rand_photon + radiate_photon3 of the reference's source on the CPU, fed the same three draws."""
    z = np.load(DEV)
    lib = azt.oracle_lib()
    lib.azo_tstat_atom.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, _fp, _fp, _fp]
    ux, uy, uz = (np.zeros(3072) for _ in range(3))
    dp = C.POINTER(C.c_double)
    lib.azo_default_uvects(ux.ctypes.data_as(dp), uy.ctypes.data_as(dp), uz.ctypes.data_as(dp))
    uv = np.stack([ux, uy, uz], 1).astype(np.float32).copy()
    mass = float(z["tstat_mass"][0])
    radiated = 0
    for row in z["tstat"]:
        v = np.array(row[0:3], dtype=np.float32)
        U = C.c_float(row[3])
        lib.azo_tstat_atom(mass, row[4], int(row[5]), int(row[6]), int(row[7]), uv.ctypes.data_as(_fp), v.ctypes.data_as(_fp), C.byref(U))
        vs = max(np.abs(row[8:11]).max(), 1e-6)
        assert np.abs(v - row[8:11]).max() <= 2e-5 * vs, row
        assert U.value == pytest.approx(row[11], rel=2e-4, abs=2e-6), row  # U is a difference of two kinetic energies
        rad = 2.73 / (4.731 - min(U.value, 0.2))
        assert rad == pytest.approx(row[12], rel=1e-5)
        radiated += row[3] + row[4] > 1e-3
    assert radiated > 30  # the radiate_photon3 branch is exercised


def test_oracle_grid_matches_reference_split_cells_bit_exactly():
    z = np.load(DEV)
    lib = azt.oracle_lib()
    for row in z["grid"]:
        cn, cs, cr = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        lib.azo_split_cells(C.c_float(row[0]), C.c_float(row[1]), C.c_float(row[2]), C.c_float(row[3]), cn, cs, cr)
        assert list(cn) == [int(v) for v in row[4:7]], row
        assert [np.float32(v) for v in cs] == [np.float32(v) for v in row[7:10]], row      # cSize, bit-exact
        assert [np.float32(v) for v in cr] == [np.float32(v) for v in row[10:13]], row     # cRevSize, bit-exact


def test_oracle_cell_pairs_match_reference_cell_dist_int():
    """Per axis: the set of (cell, neighbour, image shift) the oracle's stencil walk visits equals the set of ordered cell
    pairs cell_dist_int (src/cuCellList.cu:98-132) keeps, with the same shift sign."""
    z = np.load(DEV)
    lib = azt.oracle_lib()
    lib.azo_stencil_half.restype = C.c_int
    lib.azo_stencil_half.argtypes = [C.c_int, C.c_float, C.c_float]
    lib.azo_neighbor.restype = C.c_int
    lib.azo_neighbor.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    groups = {}
    for mx, csize, rskip, xi, xj, skip, sh, rmin in z["celldist"]:
        groups.setdefault((int(mx), float(csize), float(rskip)), {})[(int(xi), int(xj))] = (int(skip), int(sh))
    for (mx, csize, rskip), ref in groups.items():
        kept_ref = {(xi, xj): sh for (xi, xj), (skip, sh) in ref.items() if not skip}
        sten = lib.azo_stencil_half(mx, csize, rskip)
        kept = {}
        for c in range(mx):
            for o in range(-sten, sten + 1):
                cn, sh = C.c_int(), C.c_int()
                if lib.azo_neighbor(mx, c, o, C.byref(cn), C.byref(sh)):
                    assert (c, cn.value) not in kept, "a cell pair must exist once"
                    kept[(c, cn.value)] = sh.value
        # the reference's shift is +1 when the FIRST cell has the larger coordinate and the pair is wrapped; the stencil
        # walk reports the image of the neighbour: +1 when the neighbour was reached by stepping past the upper face
        assert set(kept) == set(kept_ref), (mx, csize, rskip, sorted(set(kept) ^ set(kept_ref)))
        for k, sh in kept.items():
            assert sh == kept_ref[k], (mx, csize, rskip, k, sh, kept_ref[k])


def test_oracle_wrap_and_bin_match_reference_put_periodic_and_count_cell():
    z = np.load(DEV)
    lib = azt.oracle_lib()
    lib.azo_wrap_coord.restype = C.c_float
    lib.azo_wrap_coord.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int)]
    checked = 0
    for row in z["wrapbin"]:
        xin, L, cn, cr = row[0:3], row[6:9], [int(v) for v in row[9:12]], row[12:15]
        xref, cell_ref, cross = row[15:18], int(row[18]), [int(v) for v in row[19:25]]
        x, pos_cross, neg_cross = [], [], []
        for a in range(3):
            n = C.c_int()
            revL = np.float32(1.0) / np.float32(L[a])
            x.append(lib.azo_wrap_coord(C.c_float(xin[a]), C.c_float(L[a]), C.c_float(revL), C.byref(n)))
            pos_cross.append(1 if n.value > 0 else 0)
            neg_cross.append(1 if n.value < 0 else 0)
        assert [np.float32(v) for v in x] == [np.float32(v) for v in xref], row   # wrapped coordinates, bit-exact
        assert pos_cross + neg_cross == cross
        ncell = cn[0] * cn[1] * cn[2]
        if 0 <= cell_ref < ncell:  # inside the grid the clamp of the oracle is inactive: indices must be equal
            crev = (C.c_float * 3)(*cr)
            cnum = (C.c_int * 3)(*cn)
            c = lib.azo_cell_index(C.c_float(x[0]), C.c_float(x[1]), C.c_float(x[2]), crev, cnum)
            per_axis_ok = all(0 <= int(np.floor(np.float64(np.float32(x[a])) * np.float64(np.float32(cr[a])))) < cn[a] for a in range(3))
            if per_axis_ok:
                assert c == cell_ref, row
                checked += 1
    assert checked > 300


def test_oracle_nose_hoover_recurrence_matches_reference_source():
    """before_nose + after_nose(1) of src/cuTemp.cu executed on the CPU: friction, velocity factor, kinetic energy, integral."""
    z = np.load(DEV)
    lib = azt.oracle_lib()
    lib.azo_nose_scalars.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, _fp, _fp]
    chit, ci = C.c_float(0.0), C.c_float(0.0)
    prev_te = None
    for tSt, teKin, rq, qt, ke_in, ke_out, chit_ref, ci_ref, ts_ref in z["nose"]:
        if prev_te != teKin:
            chit, ci, prev_te = C.c_float(0.0), C.c_float(0.0), teKin
        ke, ts = C.c_float(ke_in), C.c_float()
        lib.azo_nose_scalars(tSt, teKin, rq, qt, C.byref(ke), C.byref(chit), C.byref(ci), C.byref(ts))
        assert np.float32(ts.value) == np.float32(ts_ref) and np.float32(ke.value) == np.float32(ke_out)   # bit-exact
        assert np.float32(chit.value) == np.float32(chit_ref) and np.float32(ci.value) == np.float32(ci_ref)


# SURVEY 8f-4 groundwork (oracle first, no kernel yet): the reference's bond potentials and valence-angle term, cut out of
# src/cuBonds.cu / src/cuAngles.cu and executed on the CPU -> tests/golden/ref_device_bonds_angles.npz
DEV_BA = os.path.join(GOLD, "ref_device_bonds_angles.npz")


def test_oracle_bond_potentials_match_reference_device_source():
    z = np.load(DEV_BA)
    lib = azt.oracle_lib()
    lib.azo_bond_eval.restype = C.c_float
    lib.azo_bond_eval.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(C.c_float)]
    assert sorted(set(int(k) for k in z["bond"][:, 0])) == [1, 2, 3, 4, 5]
    for row in z["bond"]:
        kind, r2, r, f_ref, e_ref = int(row[0]), row[1], row[2], row[3], row[4]
        p = (C.c_float * 5)(*row[5:10])
        e = C.c_float(0.0)
        f = lib.azo_bond_eval(kind, p, r2, r, C.byref(e))
        assert f == pytest.approx(f_ref, rel=3e-6, abs=1e-9), (kind, r)
        assert e.value == pytest.approx(e_ref, rel=3e-6, abs=2e-6), (kind, r)
    # known answers: harmonic at r0 (no force, no energy), Morse at r0 (U = -C, no force)
    e = C.c_float(0.0)
    assert lib.azo_bond_eval(1, (C.c_float * 5)(25.0, 1.1, 0, 0, 0), 1.1 * 1.1, 1.1, C.byref(e)) == 0.0 and e.value == 0.0
    e = C.c_float(0.0)
    assert lib.azo_bond_eval(2, (C.c_float * 5)(4.5, 2.0, 0.96, 4.5, 0), 0.96 ** 2, 0.96, C.byref(e)) == 0.0 and e.value == pytest.approx(-4.5)


def test_oracle_valence_angle_matches_reference_device_source():
    z = np.load(DEV_BA)
    lib = azt.oracle_lib()
    fp = C.POINTER(C.c_float)
    lib.azo_angle_hcos.restype = None
    lib.azo_angle_hcos.argtypes = [C.c_float, C.c_float, fp, fp, fp, fp]
    L = z["angle_box"].astype(np.float32)
    crossed = 0
    for row in z["angle"]:
        k, cos0 = row[0], row[1]
        xyz = np.ascontiguousarray(row[2:11], dtype=np.float32)
        f_ref, e_ref = row[11:20], row[20]
        frc = np.zeros(9, dtype=np.float32)
        e = C.c_float(0.0)
        lib.azo_angle_hcos(k, cos0, xyz.ctypes.data_as(fp), L.ctypes.data_as(fp), frc.ctypes.data_as(fp), C.byref(e))
        scale = max(float(np.abs(f_ref).max()), 1e-6)
        assert np.abs(frc - f_ref).max() <= 5e-6 * scale, (k, cos0)
        assert e.value == pytest.approx(e_ref, rel=5e-6, abs=1e-9)
        assert np.abs(frc.reshape(3, 3).sum(0)).max() <= 2e-5 * scale   # the three forces cancel
        d = np.abs(xyz.reshape(3, 3)[1:] - xyz.reshape(3, 3)[0])
        crossed += int((d > 0.5 * L).any())
    assert crossed >= 5  # minimum-image branches of delta_periodic were exercised


# ---------------------------------------------------------------------------------------------------------------
# a32: host-side thermostat tables of read_tstat (src/temperature.cpp:91-259) against the reference itself:
# tests/golden/ref_tstat_tables.npz = what the reference's own read_tstat built under srand(1) (oracle/_ref/ref_serial_harness
# ... tstat, tests/golden/make_golden_decks.py)
TSTAT = os.path.join(GOLD, "ref_tstat_tables.npz")


def test_unit_vector_table_is_the_reference_table_bit_for_bit():
    """adsorb_rand_photon picks uvects[rnd % 3072] (src/cuTemp.cu:484-507): a different ORDER of the same vectors would be a
    different thermostat.  The product's built-in table (az_default_uvects, what the engine uploads when the description
    carries none) and the oracle's (azo_default_uvects) equal the 3 x 32 x 16 x 2 table read_tstat builds, double for double."""
    g = np.load(TSTAT)
    ref = g["uvect"]
    assert ref.shape == (3, abi.AZ_N_UVECT)
    dp = C.POINTER(C.c_double)
    lib = abi.load_library()  # host code: needs no GPU
    a = np.zeros((3, abi.AZ_N_UVECT))
    assert lib.az_default_uvects(a[0].ctypes.data_as(dp), a[1].ctypes.data_as(dp), a[2].ctypes.data_as(dp)) == 0
    assert np.array_equal(a, ref), int((a != ref).sum())
    o = azt.oracle_lib()
    b = np.zeros((3, abi.AZ_N_UVECT))
    o.azo_default_uvects.argtypes = [dp, dp, dp]
    o.azo_default_uvects(b[0].ctypes.data_as(dp), b[1].ctypes.data_as(dp), b[2].ctypes.data_as(dp))
    assert np.array_equal(b, ref)


def test_photon_energy_sampler_reproduces_the_reference_draw_for_draw():
    """photon_engs (src/temperature.cpp:28-89): bisection of the series CDF, tolerance 1e-3, at most 20 halvings, redraw when
    there is no sign change.  Fed with the reference's own uniform stream -- rand01() = (rand() % 10000) / 10000 of glibc under
    srand(1), the first random numbers init_md consumes -- az_photon_energy returns the reference's photon energies exactly."""
    g = np.load(TSTAT)
    ref = g["photons"]
    libc = C.CDLL("libc.so.6")
    libc.srand(int(g["srand"]))
    lib = abi.load_library()
    T = float(g["temperature"])
    out = np.zeros(len(ref))
    y = C.c_double()
    prev = 0.0
    for i in range(len(ref)):
        while True:
            x = (libc.rand() % 10000) / 10000.0
            if lib.az_photon_energy(x, T, prev, C.byref(y)):
                break
        out[i] = prev = y.value
    # atom 0 ran into the 20-halvings limit under this seed: the reference then takes "the previous value", engs[-1] -- an
    # out-of-bounds read (src/temperature.cpp:77) that returned 3.4e-320 here; every later fall-back (e.g. atoms 4 -> 5) is exact
    assert out[0] == 0.0 and abs(ref[0]) < 1e-300
    assert np.array_equal(out[1:], ref[1:]), (int((out != ref).sum()), float(np.abs(out - ref).max()))
    assert (ref[1:] == ref[:-1]).sum() >= 1  # the fall-back path is exercised inside the table as well
    # the distribution it samples: P ~ e^4 exp(-e / kT) has mean 5 kT; the loose residual tolerance (1e-3) stops the bisection
    # early and biases the mean low by ~12 % -- the reference's own behaviour, reproduced above
    assert ref.mean() == pytest.approx(5 * KB * T, rel=0.15)
