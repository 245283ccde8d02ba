"""Parity cases shared by the emulator tests (CPU, logic) and the GPU tests (the real thing).

Every case builds a seeded synthetic system, advances the CPU oracle (oracle/az_oracle.c) and an
engine behind the C ABI by the same steps and compares:
  * cell assignment, per-cell counts, first-atom offsets and the sorted permutation: bit-exact;
  * per-atom forces: |df| <= FORCE_TOL * sum_j |f_ij * delta_ij| (1e-5 relative, the north-star
    bound for fp32 after identical ordering);
  * energies / kinetic energy / thermal energy: ENERGY_TOL relative;
  * thermostat state (U, radius), velocities, positions: STATE_TOL relative to their scale.
"""
from __future__ import annotations

import os

import numpy as np

import azt
from aztotmd_b200 import abi
from aztotmd_b200.engine import AzError, Engine, Species, System

FORCE_TOL = 1e-5
ENERGY_TOL = 2e-6
STATE_TOL = 2e-5


def assert_cells_equal(do, de):
    assert do["cnum"] == de["cnum"]
    assert do["crev"] == de["crev"] and do["csize"] == de["csize"]
    assert np.array_equal(do["count"], de["count"]), "nAtInCell differs"
    assert np.array_equal(do["first"], de["first"]), "firstAtomInCell differs"
    assert np.array_equal(do["cell"], de["cell"]), "cell index of sorted atoms differs"
    assert np.array_equal(do["ids"], de["ids"]), "sorted permutation differs"


def assert_forces_close(o, do, de, tol=FORCE_TOL):
    sc = np.maximum(o.force_scale(), 1e-10)
    err = np.abs(do["frc"].astype(np.float64) - de["frc"].astype(np.float64)).max(1)
    worst = float((err / sc).max())
    assert worst <= tol, "force error %.3g of the pair-term scale (tol %.1g)" % (worst, tol)


def rel(a, b, floor=1e-12):
    return abs(a - b) / max(abs(a), abs(b), floor)


def assert_stats_close(so, se, tol=ENERGY_TOL, thermo=False, escale=None):
    floor = escale if escale is not None else max(abs(so.engKin_d), abs(so.engVdW_d), abs(so.engCoul1_d), 1e-6)
    for k in ("engKin_d", "engVdW_d", "engCoul1_d") + (("engTemp_d",) if thermo else ()):
        a, b = getattr(so, k), getattr(se, k)
        assert abs(a - b) <= tol * max(abs(a), floor), "%s: oracle %r engine %r" % (k, a, b)
    # reciprocal-space Ewald: k-space energy (fp32 phases against the oracle's double ones) and the constant term
    assert abs(so.engCoul2_d - se.engCoul2_d) <= 10 * tol * max(abs(so.engCoul2_d), floor), "engCoul2: oracle %r engine %r" % (so.engCoul2_d, se.engCoul2_d)
    assert abs(so.engCoul3 - se.engCoul3) <= tol * max(abs(so.engCoul3), floor)
    for k in ("engBond_d", "engAngle_d"):  # bonded energies (exp / acos-free forms in fp32: 5e-6 like the reference pinning)
        a, b = getattr(so, k), getattr(se, k)
        assert abs(a - b) <= max(tol, 5e-6) * max(abs(a), floor), "%s: oracle %r engine %r" % (k, a, b)
    assert abs(so.engTot - se.engTot) <= 4 * tol * max(abs(so.engTot), floor)


def run_parity(lib, system: System, nsteps=3, first_step=0, thermo=False, check_state=True, force_tol=FORCE_TOL,
               later_force_tol=None, energy_tol=ENERGY_TOL, state_tol=STATE_TOL, **engine_kw):
    """`force_tol` is the bound for the first step, where both sides start from bit-identical inputs
    (the north-star 1e-5).  On later steps the inputs themselves differ by propagated fp32 rounding
    (positions, and for surk the per-atom radii the thermostat recomputed); `later_force_tol` covers
    potentials whose pair terms cancel strongly (surk near its minimum amplifies one ulp of radius ~100x)."""
    if later_force_tol is None:
        later_force_tol = force_tol
    o = azt.Oracle(system)
    e = Engine(system, lib=lib, **engine_kw)
    try:
        e.set_option("energy_every_step", 1)
        for k in range(nsteps):
            step = first_step + k
            o.step(step, 1)
            e.step(step, 1)
            do, de = o.dump(), e.dump()
            assert_cells_equal(do, de)
            assert_forces_close(o, do, de, force_tol if k == 0 else later_force_tol)
            assert_stats_close(o.stats(), e.stats(), energy_tol, thermo=thermo)
            if check_state:
                vs = max(float(np.abs(do["vel"]).max()), 1e-6)
                assert np.abs(do["vel"] - de["vel"]).max() <= state_tol * vs, "velocities differ at step %d" % step
                assert np.abs(do["pos"] - de["pos"]).max() <= state_tol * max(system.box), "positions differ"
        if thermo:
            ao, ae = o.download(abi.ORDER_ORIGINAL), e.download(abi.ORDER_ORIGINAL)
            us = max(float(np.abs(ao["engs"]).max()), 1e-4)
            assert np.abs(ao["engs"] - ae["engs"]).max() <= 20 * state_tol * us, "thermal energies differ"
            assert np.abs(ao["radii"] - ae["radii"]).max() <= 20 * state_tol * float(np.abs(ao["radii"]).max())
        return o, e
    except Exception:
        e.close()
        raise


# ---------------------------------------------------------------------------------------- cases
def case_lj_basic(lib):
    o, e = run_parity(lib, azt.lj_system(ncell=(5, 5, 5)), nsteps=3)
    assert o.pairs_in_range() > 10000
    e.close()


def case_lj_anisotropic_even_odd_grid(lib):
    s = azt.lj_system(ncell=(6, 5, 7), seed=4)
    o, e = run_parity(lib, s, nsteps=2)
    assert tuple(e.info.cnum) == (4, 3, 4)
    e.close()


def case_lj_two_cells_per_axis_no_wrap(lib):
    # 2 cells per axis: the reference pairs cells 0 and 1 once, with no periodic image
    # (delt > mx/2 is false, src/cuCellList.cu:111); 1 cell: only intra-cell pairs
    o, e = run_parity(lib, azt.lj_system(ncell=(4, 4, 4), seed=5), nsteps=2)
    assert tuple(e.info.cnum) == (2, 2, 2)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(2, 2, 3), seed=6, cell_size=20.0), nsteps=2)
    assert tuple(e.info.cnum) == (1, 1, 1)
    e.close()


def case_lj_small_cells_wide_stencil(lib):
    s = azt.lj_system(ncell=(5, 5, 5), seed=7, cell_size=4.3)
    o, e = run_parity(lib, s, nsteps=2)
    assert tuple(e.info.cnum) == (6, 6, 6) and tuple(e.info.stencil) == (2, 2, 2)
    e.close()
    s = azt.lj_system(ncell=(5, 5, 5), seed=8, cell_size=2.9)  # 9 cells, stencil 3 (case study 2 geometry)
    o, e = run_parity(lib, s, nsteps=1)
    assert tuple(e.info.stencil) == (3, 3, 3)
    e.close()


def case_lj_many_atoms_per_cell(lib):
    # 864 atoms in 1 cell / 8 cells: several i-batches per CTA
    s = azt.lj_system(ncell=(6, 6, 6), seed=9, cell_size=40.0)
    o, e = run_parity(lib, s, nsteps=1)
    assert tuple(e.info.cnum) == (1, 1, 1)
    e.close()
    s = azt.lj_system(ncell=(6, 6, 6), seed=9, cell_size=17.0)
    o, e = run_parity(lib, s, nsteps=1)
    e.close()


def case_ionic(lib, elec, kind="buck"):
    s = azt.nacl_system(ncell=(5, 5, 5), elec=elec, kind=kind)
    o, e = run_parity(lib, s, nsteps=2)
    e.close()


def case_ewald_reciprocal(lib):
    """'elec pme': real-space + reciprocal-space + constant term on a disordered rock-salt box."""
    s = azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_EWALD, alpha=0.3, jitter=0.4, ewald_k=(6, 6, 6))
    o, e = run_parity(lib, s, nsteps=2)
    st = e.stats()
    assert st.engCoul2 > 1.0 and st.engCoul3 < -1000.0  # both terms are live
    e.close()
    s2 = azt.nacl_system(ncell=(5, 4, 6), elec=abi.ELEC_EWALD, alpha=0.3, jitter=0.4, ewald_k=(5, 4, 7))  # anisotropic box and k ranges
    o, e = run_parity(lib, s2, nsteps=2)
    e.close()
    # the list path with a rebuild interval runs the same k-space passes after the list force pass
    s3 = azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_EWALD, alpha=0.3, ewald_k=(6, 6, 6))
    _rebuild_vs_every_step(lib, s3, 1.0, 0, nsteps=6)


def case_ionic_efield_frozen(lib):
    s = azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_DIRECT)
    s.efield = (0.01, 0.0, -0.02)
    s.species[1].frozen = 1
    o, e = run_parity(lib, s, nsteps=2)
    so, se = o.stats(), e.stats()
    assert so.engElecField != 0.0
    assert abs(so.engElecField - se.engElecField) <= 1e-5 * abs(so.engElecField)
    # frozen chlorines did not move
    a0 = np.asarray(s.pos, dtype=np.float32)
    a1 = e.download(abi.ORDER_ORIGINAL)["pos"]
    cl = np.asarray(s.types) == 1
    assert np.array_equal(a1[cl].astype(np.float32), a0[cl])
    e.close()


def case_surk_radiative(lib):
    # case study 2 in miniature: surk + radiative thermostat; equilibration scaling on steps 2 and 4
    s = azt.surk_system(n_side=6, n_equil=4, eq_freq=2)
    o, e = run_parity(lib, s, nsteps=5, thermo=True, state_tol=1e-4, later_force_tol=1e-4)
    st = e.stats()
    assert st.engTemp > 0.0
    e.close()


def case_lj_radiative(lib):
    # case study 1 flavour: LJ + radiative thermostat, zero initial velocities
    s = azt.lj_system(ncell=(5, 5, 5), seed=11, vel=False, tstat=abi.TSTAT_RADI, temperature=298.0)
    s.species[0].radA, s.species[0].radB, s.species[0].mxEng = 2.73, 4.731, 0.2
    o, e = run_parity(lib, s, nsteps=3, thermo=True, state_tol=1e-4)
    e.close()


def case_equilibration_scaling(lib):
    s = azt.lj_system(ncell=(5, 5, 5), seed=12, n_equil=10, eq_freq=1, temperature=120.0)
    o, e = run_parity(lib, s, nsteps=3, first_step=1)
    st = e.stats()
    assert st.engKin == np.float32(s.t_kin)  # after_tscale: engKin := teKin (src/cuTemp.cu:109-113)
    e.close()


def case_face_crossings_and_pressure(lib):
    # hot, fast atoms: many face crossings in 24 steps -> counters, cumulative momenta, pressure ring
    s = azt.lj_system(ncell=(5, 5, 5), seed=13, temperature=3000.0)
    s.timestep = 0.004
    a = s.box[0] / 5
    s.pos = (s.pos + 0.25 * a - 0.1) % np.asarray(s.box)  # park a lattice plane 0.1 A inside each upper face
    o = azt.Oracle(s)
    e = Engine(s, lib=lib)
    o.step(0, 24); e.step(0, 24)
    so, se = o.stats(), e.stats()
    assert sum(so.specAcBoxPos[0]) + sum(so.specAcBoxNeg[0]) > 20
    # trajectories agree to rounding for these few steps, so the integer counters must be equal
    assert list(so.specAcBoxPos[0]) == list(se.specAcBoxPos[0]) and list(so.specAcBoxNeg[0]) == list(se.specAcBoxNeg[0])
    for a in range(3):
        assert abs(so.posMom[a] - se.posMom[a]) <= 1e-4 * max(abs(so.posMom[a]), 1e-3)
        assert abs(so.negMom[a] - se.negMom[a]) <= 1e-4 * max(abs(so.negMom[a]), 1e-3)
    assert so.pressure != 0.0 and abs(so.pressure - se.pressure) <= 1e-3 * abs(so.pressure)
    e.close()


def _rebuild_vs_every_step(lib, s, skin, variant, nsteps=8, thermo=False, tol=1e-5):
    """Extension (BASELINE config 3): binning every R steps with a skin gives the same forces and
    energies as binning every step.  variant 0 = Verlet list between rebuilds, 1 = stale cell tiles."""
    e1 = Engine(s, lib=lib)
    e2 = Engine(s, lib=lib, rebuild_every=5, skin=skin, pair_variant=variant)
    o = azt.Oracle(s)
    for eng in (e1, e2):
        eng.set_option("energy_every_step", 1)
    for step in range(nsteps):
        e1.step(step, 1); e2.step(step, 1); o.step(step, 1)
        s1, s2, so = e1.stats(), e2.stats(), o.stats()
        sc = max(abs(s1.engVdW_d), abs(s1.engCoul1_d), abs(s1.engKin_d))
        assert abs(s1.engVdW_d - s2.engVdW_d) <= tol * sc and abs(s1.engCoul1_d - s2.engCoul1_d) <= tol * sc
        assert abs(s1.engKin_d - s2.engKin_d) <= tol * sc
        assert s1.n_pairs == s2.n_pairs == o.pairs_in_range(), (s1.n_pairs, s2.n_pairs, o.pairs_in_range())
        if thermo:
            assert abs(s1.engTemp_d - s2.engTemp_d) <= 1e-4 * max(abs(s1.engTemp_d), 1e-6)
    a1, a2 = e1.download(abi.ORDER_ORIGINAL), e2.download(abi.ORDER_ORIGINAL)
    assert np.abs(a1["pos"] - a2["pos"]).max() < 1e-3 and np.abs(a1["vel"] - a2["vel"]).max() < 1e-3 * np.abs(a1["vel"]).max()
    ao = o.download(abi.ORDER_ORIGINAL)
    assert np.abs(ao["pos"] - a2["pos"]).max() < 1e-3
    # forces of the last step, atom by atom (original order), against the oracle
    d2, do = azt.sort_by_id(e2.dump()), azt.sort_by_id(o.dump())
    fs = np.zeros(o.n, dtype=np.float32); fs[o.dump()["ids"]] = o.force_scale()
    assert (np.abs(d2["frc"] - do["frc"]).max(1) <= 5e-4 * np.maximum(fs, 1e-10)).all()  # trajectories 8 steps apart in rounding
    e1.close(); e2.close()


def case_rebuild_interval_list(lib):
    _rebuild_vs_every_step(lib, azt.lj_system(ncell=(6, 6, 6), seed=14, cell_size=9.3), 0.8, 0)


def case_rebuild_interval_list_wide_stencils(lib):
    # half- and third-size cells (stencil 2 and 3): the list build prunes cells by the atom's own position and walks
    # z-intervals of the columns; every in-range pair must still be found (pair counts equal the oracle's)
    _rebuild_vs_every_step(lib, azt.lj_system(ncell=(7, 7, 7), seed=17, cell_size=4.7), 0.6, 0, nsteps=5)
    _rebuild_vs_every_step(lib, azt.lj_system(ncell=(6, 7, 8), seed=18, cell_size=3.2), 0.5, 0, nsteps=4)
    # the generic cell walk (tiny grids use it) gives the same list as the fast one
    s = azt.lj_system(ncell=(7, 7, 7), seed=17, cell_size=4.7)
    a, b = Engine(s, lib=lib, rebuild_every=4, skin=0.6), Engine(s, lib=lib, rebuild_every=4, skin=0.6)
    b.set_option("list_fast", 0)
    a.step(0, 6); b.step(0, 6)
    da, db = a.dump(), b.dump()
    assert np.array_equal(da["pos"], db["pos"]) and np.abs(da["frc"] - db["frc"]).max() <= 1e-5 * np.abs(db["frc"]).max()
    assert a.stats().n_pairs == b.stats().n_pairs
    a.close(); b.close()


def case_rebuild_interval_stale_tiles(lib):
    _rebuild_vs_every_step(lib, azt.lj_system(ncell=(6, 6, 6), seed=14, cell_size=9.3), 0.8, 1)


def case_rebuild_interval_list_ionic(lib):
    s = azt.nacl_system(ncell=(5, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.6)  # 3 cells of 9.83 >= rc + skin
    _rebuild_vs_every_step(lib, s, 0.6, 0, nsteps=6)


def case_rebuild_interval_list_surk_radiative(lib):
    s = azt.surk_system(n_side=6, cell_size=3.3)
    _rebuild_vs_every_step(lib, s, 0.5, 0, nsteps=6, thermo=True, tol=1e-4)


def case_list_capacity_is_reported(lib):
    s = azt.lj_system(ncell=(6, 6, 6), seed=14, cell_size=9.3)
    e = Engine(s, lib=lib, rebuild_every=5, skin=0.8, max_neigh=20)
    e.step(0, 1)
    try:
        e.stats()
    except AzError as err:
        assert err.status == abi.AZ_ERR_OVERFLOW and "max_neigh" in str(err)
    else:
        raise AssertionError("neighbour-list overflow not reported")
    e.close()


def case_skin_violation_is_reported(lib):
    s = azt.lj_system(ncell=(6, 6, 6), seed=15, cell_size=9.3, temperature=5000.0)
    s.timestep = 0.005
    e = Engine(s, lib=lib, rebuild_every=50, skin=0.05)
    e.step(0, 20)
    try:
        e.stats()
    except AzError as err:
        assert err.status == abi.AZ_ERR_OVERFLOW
    else:
        raise AssertionError("skin violation not reported")
    e.close()


def case_download_upload_roundtrip(lib):
    s = azt.lj_system(ncell=(5, 5, 5), seed=16)
    o = azt.Oracle(s)
    e = Engine(s, lib=lib)
    o.step(0, 2); e.step(0, 2)
    for order in (abi.ORDER_SORTED, abi.ORDER_ORIGINAL):
        ao, ae = o.download(order), e.download(order)
        assert np.array_equal(ao["ids"], ae["ids"]) and np.array_equal(ao["types"], ae["types"])
        assert np.abs(ao["pos"] - ae["pos"]).max() <= 1e-5 * max(s.box)
    ae = e.download(abi.ORDER_ORIGINAL)
    assert np.array_equal(ae["ids"], np.arange(len(s.pos)))
    # upload the initial state again: the next step must reproduce step 0 of a fresh engine
    e.upload(np.asarray(s.pos), np.asarray(s.vel))
    e.step(0, 1)
    f = Engine(s, lib=lib)
    f.step(0, 1)
    d1, d2 = azt.sort_by_id(e.dump()), azt.sort_by_id(f.dump())
    assert np.array_equal(d1["pos"], d2["pos"]) and np.array_equal(d1["cell"], d2["cell"])
    assert np.abs(d1["frc"] - d2["frc"]).max() <= 1e-6 * np.abs(d2["frc"]).max()
    # pipelined uploads (double-buffered staging): upload k+1 is issued while step k is still in flight
    g = Engine(s, lib=lib)
    pos0, vel0 = np.asarray(s.pos), np.asarray(s.vel)
    g.upload(pos0, vel0, asynchronous=True)
    for k in range(3):
        g.step(0, 1)
        g.upload(pos0, vel0, asynchronous=True)
    g.step(0, 1)
    d3 = azt.sort_by_id(g.dump())
    assert np.array_equal(d3["pos"], d2["pos"]) and np.array_equal(d3["frc"], d2["frc"])
    # float entry points: az_engine_upload_f32_async narrows nothing the double path would not (bitwise the same step), and
    # az_engine_download_f32 hands back what az_engine_download does, in atoms.xyz order
    g.upload_f32(pos0, vel0)
    g.step(0, 1)
    d4 = azt.sort_by_id(g.dump())
    assert np.array_equal(d4["pos"], d2["pos"]) and np.array_equal(d4["frc"], d2["frc"])
    a64, a32 = g.download(abi.ORDER_ORIGINAL), g.download_f32()
    assert np.array_equal(a32["pos"], a64["pos"].astype(np.float32)) and np.array_equal(a32["vel"], a64["vel"].astype(np.float32))
    e.close(); f.close(); g.close()


def case_determinism(lib):
    """No float atomics on the force path: two runs are bitwise identical."""
    s = azt.nacl_system(ncell=(5, 5, 5))
    outs = []
    for _ in range(2):
        e = Engine(s, lib=lib)
        e.step(0, 3)
        d = e.dump()
        outs.append((d["frc"].copy(), d["pos"].copy(), d["vel"].copy(), e.stats().engVdW_d))
        e.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]


def case_fused_kick_drift_is_bitwise_the_separate_kernels(lib):
    """Inside one az_engine_step call the second half kick of step n and the kick + drift of step n+1 share
    a pass over the atoms; the trajectory must be bitwise the one of single-step calls (separate kernels)."""
    systems = [(azt.lj_system(ncell=(6, 6, 6), seed=31, temperature=800.0), {}),
               (azt.lj_system(ncell=(6, 6, 6), seed=32, cell_size=9.3), dict(rebuild_every=4, skin=0.8)),   # integrator rides on the list kernel
               (azt.lj_system(ncell=(6, 6, 6), seed=34, temperature=600.0), dict(pair_variant=2)),            # ... and bins the next step
               (azt.nacl_system(ncell=(5, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.6), dict(rebuild_every=3, skin=0.6)),
               (azt.surk_system(n_side=6, n_equil=4, eq_freq=2), {}),
               (azt.lj_system(ncell=(5, 5, 5), seed=33, vel=False, tstat=abi.TSTAT_RADI, temperature=298.0), dict(rebuild_every=3, skin=0.6, _rad=True))]
    for s, kw in systems:
        kw = dict(kw)
        if kw.pop("_rad", False):
            s.species[0].radA, s.species[0].radB, s.species[0].mxEng = 2.73, 4.731, 0.2
        a, b = Engine(s, lib=lib, **kw), Engine(s, lib=lib, **kw)
        a.step(0, 11)
        for k in range(11):
            b.step(k, 1)
        da, db = a.dump(), b.dump()
        for key in ("pos", "vel", "frc", "ids", "cell"):
            assert np.array_equal(da[key], db[key]), key
        sa, sb = a.stats(), b.stats()
        for key in ("engKin_d", "engVdW_d", "engTemp_d", "n_pairs", "pressure"):
            assert getattr(sa, key) == getattr(sb, key), key
        assert list(sa.posMom) == list(sb.posMom) and list(sa.specAcBoxPos[0]) == list(sb.specAcBoxPos[0])
        c = Engine(s, lib=lib, **kw)
        c.set_option("fuse_integrate", 0)
        c.step(0, 11)
        assert np.array_equal(c.dump()["pos"], da["pos"])
        a.close(); b.close(); c.close()


def case_graph_replay_is_bitwise_the_plain_launches(lib):
    """az_cuda_opts.use_graph: a call is captured once per shape and replayed; the kernels read the step from the device
    (thermostat RNG counter, pressure ring).  Trajectories, thermostat state and statistics must be BITWISE those of the
    plain launches, over repeated calls of the same and of different lengths, on the every-step schedule with the
    radiative thermostat and on the list path with a rebuild interval and the riding integrator.  (On the host emulator
    there is no graph API: the same relative-step kernels run as plain launches.)"""
    systems = [(azt.surk_system(n_side=5, spacing=1.3, box=30.0), {}),
               (azt.lj_system(ncell=(6, 6, 6), seed=95, cell_size=9.3, n_equil=6, eq_freq=3), dict(rebuild_every=4, skin=0.8)),
               (azt.lj_system(ncell=(5, 5, 5), seed=96, tstat=abi.TSTAT_RADI), {})]
    for s, kw in systems:
        a, b = Engine(s, lib=lib, **kw), Engine(s, lib=lib, use_graph=True, **kw)
        step = 0
        for n in (3, 4, 4, 4, 1, 4, 7, 7):   # repeated shapes are replays
            a.step(step, n); b.step(step, n)
            step += n
            da, db = a.dump(), b.dump()
            assert np.array_equal(da["pos"], db["pos"]) and np.array_equal(da["vel"], db["vel"]) and np.array_equal(da["frc"], db["frc"]), step
            sa, sb = a.stats(), b.stats()
            for k in ("engKin_d", "engVdW_d", "engTemp_d", "engTot", "pressure", "n_pairs", "step"):
                assert getattr(sa, k) == getattr(sb, k), (k, step, getattr(sa, k), getattr(sb, k))
        ta, tb = a.download(abi.ORDER_ORIGINAL), b.download(abi.ORDER_ORIGINAL)
        assert np.array_equal(ta["engs"], tb["engs"], equal_nan=True) and np.array_equal(ta["radii"], tb["radii"], equal_nan=True)
        a.close(); b.close()


def case_tile_list_is_bitwise_the_global_rows(lib):
    """list_format 2 (tile lists: 16-bit rows into a shared-memory copy of each tile's neighbourhood, az_tlist.cuh) against
    list_format 1 (32-bit rows gathered from global memory): same candidate ranges, same order, same pair arithmetic -- the
    trajectories must agree bit for bit, on every pair-kernel instantiation, with rows through periodic faces, with rows
    that outgrow the build's shared columns, and when the footprint is the whole (small) system."""
    systems = [(azt.lj_system(ncell=(6, 6, 6), seed=32, cell_size=9.3), dict(rebuild_every=4, skin=0.8)),     # integrator rides on the list kernel
               (azt.lj_system(ncell=(8, 6, 7), seed=35, temperature=500.0), dict(rebuild_every=5, skin=1.0)),  # anisotropic grid
               (azt.lj_system(ncell=(6, 6, 6), seed=34, temperature=600.0), dict(pair_variant=2)),              # rebuilt every step
               (azt.lj_system(ncell=(5, 5, 5), seed=36, cell_size=4.5), dict(rebuild_every=3, skin=0.5)),       # wide stencil
               (azt.nacl_system(ncell=(5, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.6), dict(rebuild_every=3, skin=0.6)),
               (azt.surk_system(n_side=6, n_equil=4, eq_freq=2), dict(rebuild_every=3, skin=0.5)),
               (azt.lj_system(ncell=(5, 5, 5), seed=33, vel=False, tstat=abi.TSTAT_RADI, temperature=298.0), dict(rebuild_every=3, skin=0.6, _rad=True))]
    # (the emulator build runs the plain Lennard-Jones rows of list_format 1 through the packed f32x2 form, which keeps two
    # partial sums per component: there the comparison is to rounding, everywhere else -- and always on the GPU -- bitwise)
    emu = "emu" in os.path.basename(getattr(lib, "_name", ""))
    for s, kw in systems:
        kw = dict(kw)
        rad = kw.pop("_rad", False)
        if rad:
            s.species[0].radA, s.species[0].radB, s.species[0].mxEng = 2.73, 4.731, 0.2
        exact = not (emu and len(s.species) == 1 and s.elec_kind == abi.ELEC_NONE and all(v["kind"] == abi.VDW_LJ for v in s.vdw.values()))
        a, b = Engine(s, lib=lib, list_format=1, **kw), Engine(s, lib=lib, list_format=2, **kw)
        assert a.info.tile_list == 0 and b.info.tile_list == 1
        a.step(0, 7); b.step(0, 7)
        da, db = a.dump(), b.dump()
        for key in ("pos", "vel", "frc", "ids", "cell"):
            if exact or key in ("ids", "cell"):
                assert np.array_equal(da[key], db[key]), key
            else:
                assert np.abs(da[key] - db[key]).max() <= 1e-5 * np.abs(da[key]).max(), key
        sa, sb = a.stats(), b.stats()
        for key in ("engKin_d", "engVdW_d", "engCoul1_d", "engTemp_d", "n_pairs", "pressure"):
            if exact or key == "n_pairs":
                assert getattr(sa, key) == getattr(sb, key), key
            else:
                assert abs(getattr(sa, key) - getattr(sb, key)) <= 2e-6 * max(1e-30, abs(getattr(sa, key))), key
        b.refresh_info()
        assert 0 < b.info.tile_atoms_seen <= b.info.tile_atoms and 0 < b.info.tile_runs_seen <= 64
        # single-step calls; the switch by name takes an engine created for tile lists back to the global rows
        c = Engine(s, lib=lib, list_format=2, **kw)
        for k in range(7):
            c.step(k, 1)
        assert np.array_equal(c.dump()["pos"], db["pos"]) and np.array_equal(c.dump()["frc"], db["frc"])
        c.close()
        c = Engine(s, lib=lib, list_format=2, **kw)
        c.set_option("tile_list", 0)
        c.step(0, 7)
        assert np.array_equal(c.dump()["pos"], da["pos"]) and np.array_equal(c.dump()["frc"], da["frc"])
        a.close(); b.close(); c.close()
    # a footprint allotment that is too small is reported, not silently wrong
    s = azt.lj_system(ncell=(12, 12, 12), seed=37)
    e = Engine(s, lib=lib, rebuild_every=4, skin=0.8, list_format=2, tile_atoms=300)
    e.step(0, 2)
    try:
        e.stats()
    except AzError as err:
        assert err.status == abi.AZ_ERR_OVERFLOW and "tile_atoms" in str(err)
    else:
        raise AssertionError("footprint overflow not reported")
    e.close()


def case_tile_kernel_variant(lib):
    """pair_variant = 1: the shared-memory cell-tile sweep (one thread per atom) instead of the default
    warp-per-atom sweep -- same parity bars."""
    o, e = run_parity(lib, azt.lj_system(ncell=(6, 6, 6), seed=61), nsteps=2, pair_variant=1)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(5, 5, 5), seed=62, cell_size=2.9), nsteps=1, pair_variant=1)
    e.close()
    o, e = run_parity(lib, azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_FENNEL), nsteps=2, pair_variant=1)
    e.close()
    o, e = run_parity(lib, azt.surk_system(n_side=6, n_equil=4, eq_freq=2), nsteps=3, thermo=True, state_tol=1e-4, later_force_tol=1e-4, pair_variant=1)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(6, 6, 6), seed=9, cell_size=17.0), nsteps=1, pair_variant=1)  # several i-batches per CTA
    e.close()


def case_thread_kernel_variant(lib):
    """pair_variant = 4: the cell sweep with one thread per atom (az_k_pair_thread), what systems with cells much wider than
    the cut-off use automatically (case study 1 of the reference) -- same parity bars, with and without the per-atom
    pruning of stencil cells, on every pair-kernel instantiation."""
    o, e = run_parity(lib, azt.lj_system(ncell=(6, 6, 6), seed=63), nsteps=2, pair_variant=4)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(6, 6, 6), seed=64, cell_size=17.0), nsteps=3, pair_variant=4)   # wide cells: pruning drops most of the stencil
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(5, 5, 5), seed=65, cell_size=2.9), nsteps=1, pair_variant=4)     # wide stencil
    e.close()
    o, e = run_parity(lib, azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_FENNEL), nsteps=2, pair_variant=4)
    e.close()
    o, e = run_parity(lib, azt.surk_system(n_side=6, n_equil=4, eq_freq=2), nsteps=3, thermo=True, state_tol=1e-4, later_force_tol=1e-4, pair_variant=4)
    e.close()
    # the same trajectory as the warp-per-atom sweep (to rounding: the summation order differs), with the pruning switched off too
    s = azt.lj_system(ncell=(6, 6, 6), seed=66, cell_size=20.0, temperature=300.0)
    a, b, c = Engine(s, lib=lib, pair_variant=3), Engine(s, lib=lib, pair_variant=4), Engine(s, lib=lib, pair_variant=4)
    c.set_option("warp_prune", 0)
    for x in (a, b, c):
        x.step(0, 4)
    da, db, dc = a.dump(), b.dump(), c.dump()
    assert np.array_equal(da["ids"], db["ids"]) and np.array_equal(da["cell"], db["cell"])
    scale = np.abs(da["frc"]).max()
    assert np.abs(da["frc"] - db["frc"]).max() <= 1e-5 * scale and np.abs(dc["frc"] - db["frc"]).max() <= 1e-5 * scale
    assert a.stats().n_pairs == b.stats().n_pairs == c.stats().n_pairs
    a.close(); b.close(); c.close()


def case_stencil_table_is_bitwise_the_per_atom_stencil(lib):
    """The warp-per-atom sweep reads the stencil of a cell from a table built at create time (az_k_stencil_table) instead of
    working it out for every atom: same cells, same order, same merged ranges -- bitwise the same trajectory, with pruning
    (wide cells), wide stencils, periodic images in tiny grids, and species tables."""
    systems = [azt.lj_system(ncell=(6, 6, 6), seed=71), azt.lj_system(ncell=(6, 6, 6), seed=72, cell_size=17.0),
               azt.lj_system(ncell=(5, 5, 5), seed=73, cell_size=2.9), azt.lj_system(ncell=(3, 3, 3), seed=74),
               azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_FENNEL), azt.surk_system(n_side=6, n_equil=4, eq_freq=2)]
    for s in systems:
        a, b = Engine(s, lib=lib, pair_variant=3), Engine(s, lib=lib, pair_variant=3)
        b.set_option("stencil_table", 0)
        a.step(0, 4); b.step(0, 4)
        da, db = a.dump(), b.dump()
        for key in ("pos", "vel", "frc", "ids", "cell"):
            assert np.array_equal(da[key], db[key]), key
        assert a.stats().n_pairs == b.stats().n_pairs and a.stats().engVdW_d == b.stats().engVdW_d
        a.close(); b.close()


def case_list_every_step_variant(lib):
    """pair_variant = 2 on the every-step schedule (what large systems use automatically): the Verlet list is
    rebuilt every step with radius rc, so the recorded set IS the in-range set -- same parity bars."""
    o, e = run_parity(lib, azt.lj_system(ncell=(6, 6, 6), seed=63), nsteps=3, pair_variant=2)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(5, 5, 5), seed=64, cell_size=2.9), nsteps=2, pair_variant=2)
    e.close()
    o, e = run_parity(lib, azt.nacl_system(ncell=(5, 5, 5), elec=abi.ELEC_DIRECT), nsteps=2, pair_variant=2)
    e.close()
    o, e = run_parity(lib, azt.surk_system(n_side=6, n_equil=4, eq_freq=2), nsteps=3, thermo=True, state_tol=1e-4, later_force_tol=1e-4, pair_variant=2)
    e.close()
    o, e = run_parity(lib, azt.lj_system(ncell=(4, 4, 4), seed=5), nsteps=2, pair_variant=2)  # 2 cells per axis: generic cell walk
    e.close()


def case_force_autocap_fires(lib):
    """force_autocap (src/cuPairs.cu:94-115): two pairs are put 0.5 A and 0.9 A apart, where the Lennard-Jones F/r is
    1.8e10 and 1.6e7 -- far beyond MX_FRC = 1e4 -- so the clamp decides the force on those four atoms.  Every pair pass
    is driven through it: the warp sweep (default for small systems), the shared-memory tiles, the Verlet list rebuilt
    every step (whose packed f32x2 loop leaves the clamp to a scalar redo of the atoms that need it) and the list
    reused between rebuilds."""
    s = azt.lj_system(ncell=(5, 5, 5), seed=71, vel=False)
    pos = s.pos.copy()
    pos[1] = pos[0] + np.array([0.3, -0.4, 0.0])          # r = 0.5 A
    pos[101] = pos[100] + np.array([0.0, 0.54, 0.72])     # r = 0.9 A
    s.pos = pos
    for kw in ({}, {"pair_variant": 1}, {"pair_variant": 2}, {"rebuild_every": 4, "skin": 0.6}):
        if "rebuild_every" in kw:
            s.cell_size = 2.5 * azt.AR_SIGMA + 0.6
        o, e = run_parity(lib, s, nsteps=1, **kw) # zero velocities, zero initial forces: step 0 evaluates the initial positions
        f = azt.sort_by_id(e.dump())["frc"].astype(np.float64)
        # the clamped pair term is exactly MX_FRC * delta; the other ~55 neighbours add forces of order 0.1
        assert abs(np.linalg.norm(f[0]) - 1e4 * 0.5) < 5.0 and abs(np.linalg.norm(f[1]) - 1e4 * 0.5) < 5.0, (kw, f[0], f[1])
        assert abs(np.linalg.norm(f[100]) - 1e4 * 0.9) < 5.0 and abs(np.linalg.norm(f[101]) - 1e4 * 0.9) < 5.0, (kw, f[100], f[101])
        e.close()


def bonded_lj_system(seed=81, angles=True, **kw):
    """500 argon atoms with a list of constant bonds (harmonic and Morse alternating; the last one reaches through a
    periodic face) and harmonic-cosine valence angles -- the layout of tests/golden/make_golden.py's bonded decks."""
    s = azt.lj_system(ncell=(5, 5, 5), seed=seed, **kw)
    n = len(s.pos)
    s.bond_types = [("harm", [5.0, 3.8]), ("mors", [0.4, 1.5, 4.0, 0.4]), ("pdn", [0.3, 1.2, 4.1, 0.3, 50.0]), ("buck", [900.0, 0.35, 60.0]),
                    ("e612", [800.0, 0.36, 50.0, 20.0, 10.0])]
    b = [(4 * k, 4 * k + 1, 1 + (k % 5)) for k in range(0, n // 4, 3)] + [(2, 4 * 5 * 5 * 4 + 2, 1)]
    s.bonds = np.array(b, dtype=np.int32)
    if angles:
        s.angle_types = [(6.0, -0.3), (2.5, 0.45)]
        a = [(4 * k, 4 * k + 1, 4 * k + 2, 1 + (k % 2)) for k in range(1, n // 4, 4)] + [(1, 3, 4 * 5 * 4 + 1, 2)]
        s.angles = np.array(a, dtype=np.int32)
    return s


def case_bonds_and_angles(lib):
    """constant bonds (all five potentials) and valence angles on top of the LJ pair forces, followed through
    the sort of every step: forces 1e-5, engBond / engAngle 5e-6, on the warp sweep, the tile sweep, the list rebuilt
    every step, and the list reused between rebuilds (where the integrator may not ride on the pair pass)."""
    for kw in ({}, {"pair_variant": 1}, {"pair_variant": 2}):
        o, e = run_parity(lib, bonded_lj_system(), nsteps=4, **kw)
        st = e.stats()
        assert abs(st.engBond_d) > 1e-2 and abs(st.engAngle_d) > 1e-3 and abs(st.engTot - (st.engKin + st.engVdW + st.engBond + st.engAngle)) < 1e-3
        e.close()
    o, e = run_parity(lib, bonded_lj_system(angles=False, cell_size=2.5 * azt.AR_SIGMA + 0.6), nsteps=6, rebuild_every=3, skin=0.6)
    e.close()
    # degrees of freedom: every bond removes one (src/sys_init.cpp:1099-1102)
    s = bonded_lj_system()
    assert s.deg_free == 3 * len(s.pos) - len(s.bonds)


def case_bonds_match_the_reference_serial_build(lib):
    """The engine on the reference-generated decks with bonds / bonds + angles (tests/golden/ref_serial_lj5_bonds*.npz: forces and
    bonded energies of the reference's own serial build, exec_bondlist / exec_anglelist): first force evaluation from the
    fixture's positions, 1e-5 of the per-atom term scale, energies 5e-6."""
    import ast
    import os
    names = {v: k for k, v in abi.BOND_BY_NAME.items()}
    for fx in ("ref_serial_lj5_bonds.npz", "ref_serial_lj5_bonds_angles.npz"):
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fx))
        meta = ast.literal_eval(str(z["meta"]))
        s = System(pos=z["pos"], types=z["types"].astype(np.int32), box=tuple(meta["box"]), species=[Species("Ar", azt.AR_MASS)],
                   timestep=0.001, cell_size=meta["rc"], temperature=86.6)
        s.set_vdw(0, 0, "lnjs", meta["rc"], [azt.AR_EPS, azt.AR_SIGMA])
        s.bond_types = [(names[kind], list(p)) for kind, p in meta["bond_types"]]
        s.bonds = np.array(meta["bonds"], dtype=np.int32)
        if "angles" in meta:
            s.angle_types = [tuple(t) for t in meta["angle_types"]]
            s.angles = np.array(meta["angles"], dtype=np.int32)
        o, e = run_parity(lib, s, nsteps=1)  # zero velocities and zero initial forces: step 0 evaluates the fixture's positions
        d = azt.sort_by_id(e.dump())
        scale = np.zeros(o.n, dtype=np.float32); scale[o.dump()["ids"]] = o.force_scale()
        err = np.abs(d["frc"].astype(np.float64) - z["frc"]).max(1)
        assert (err <= 1e-5 * np.maximum(scale, 1e-12)).all(), float((err / np.maximum(scale, 1e-12)).max())
        st = e.stats()
        assert abs(st.engBond_d - float(z["engBond"])) <= 5e-6 * abs(float(z["engBond"]))
        if "angles" in meta:
            assert abs(st.engAngle_d - float(z["engAngle"])) <= 5e-6 * abs(float(z["engAngle"]))
        e.close()


def case_nose_hoover(lib):
    """'temperature T nose tau' (SURVEY 8f-2): the friction is advanced twice per step (before verlet_1stage and in
    apply_tstat, src/main.cu:291-299, src/cuTemp.cu:787-794); hot start well above the target so that it acts visibly."""
    s = azt.lj_system(ncell=(5, 5, 5), seed=81, temperature=300.0, tstat=abi.TSTAT_NOSE)
    s.temperature = 120.0   # target below the initial kinetic temperature
    s.tstat_tau = 0.02
    o, e = run_parity(lib, s, nsteps=6, state_tol=5e-5)
    so, se = o.stats(), e.stats()
    ke0 = 1.5 * len(s.pos) * 300.0 * 8.617e-5
    assert se.engKin_d < 0.9 * ke0            # the thermostat removed kinetic energy
    e.close()
    # ... together with the equilibration rescale and a multi-step call
    s = azt.lj_system(ncell=(5, 5, 5), seed=82, temperature=150.0, tstat=abi.TSTAT_NOSE, n_equil=4, eq_freq=2)
    s.tstat_tau = 0.05
    o = azt.Oracle(s)
    e = Engine(s, lib=lib, rebuild_every=3, skin=0.6)
    o.step(0, 7); e.step(0, 7)
    so, se = o.stats(), e.stats()
    assert abs(so.engKin_d - se.engKin_d) <= 2e-5 * so.engKin_d and abs(so.engVdW_d - se.engVdW_d) <= 2e-5 * abs(so.engVdW_d)
    e.close()


def case_reset_vels(lib):
    """'reset_vels n': every n-th step the velocities are zeroed instead of the second half kick (quenching)."""
    s = azt.lj_system(ncell=(5, 5, 5), seed=83, temperature=200.0)
    s.reset_vels = 3
    o, e = run_parity(lib, s, nsteps=7)
    e.close()
    s = azt.surk_system(n_side=6)   # the radiative thermostat still acts on the zeroed velocities (apply_tstat follows zero_vel)
    s.reset_vels = 2
    o, e = run_parity(lib, s, nsteps=4, thermo=True, state_tol=1e-4, later_force_tol=1e-4)
    e.close()


def case_rdf_from_cells_equals_brute_force(lib):
    """the cell-based RDF histogram holds exactly the pair counts of the reference's O(N^2) brute_rdf
    (integers, so equality is exact), for one and two species, radii below and above the force cut-off, stale cells."""
    cases = [(azt.lj_system(ncell=(9, 9, 9), seed=91), {}, 11.0, 0.05),                       # rRDF > rc: two cell layers
             (azt.lj_system(ncell=(6, 6, 6), seed=92, cell_size=9.3), dict(rebuild_every=4, skin=0.8), 7.0, 0.1),  # stale cells
             (azt.nacl_system(ncell=(5, 5, 5)), {}, 8.0, 0.02),                               # three species pairs
             (azt.surk_system(n_side=6, spacing=1.3, box=35.0), {}, 8.0, 0.02)]               # case-study-2 geometry, 7 cells wide stencil
    for s, kw, r_max, dr in cases:
        o = azt.Oracle(s)
        e = Engine(s, lib=lib, **kw)
        e.rdf_begin(r_max, dr)
        total = None
        for k in range(3):
            o.step(k, 1); e.step(k, 1)
            e.rdf_sample()
            counts, ns = e.rdf_counts()
            ref = o.rdf_counts(r_max, dr, counts.shape[0], counts.shape[1])
            total = ref if total is None else total + ref
            assert ns == k + 1
            if k == 0:
                assert np.array_equal(counts, ref), (int(np.abs(counts.astype(np.int64) - ref.astype(np.int64)).sum()), int(ref.sum()))
        # later steps: positions agree to rounding, a pair exactly on a bin edge may move by one bin
        diff = np.abs(counts.astype(np.int64) - total.astype(np.int64))
        assert diff.sum() <= 1e-4 * total.sum() + 4 and counts.sum() > 1000
        assert abs(int(counts.sum()) - int(total.sum())) <= 2
        e.close()
    # a radius that would need more than one image of a cell is refused, not silently wrong
    s = azt.lj_system(ncell=(4, 4, 4), seed=93)
    e = Engine(s, lib=lib)
    try:
        e.rdf_begin(10.0, 0.1)
    except AzError as err:
        assert err.status == abi.AZ_ERR_UNSUPPORTED
    else:
        raise AssertionError("expected AZ_ERR_UNSUPPORTED")
    e.close()


def case_error_paths(lib):
    s = azt.lj_system(ncell=(3, 3, 3))
    s.cell_size = 0.0
    for mutate, status in ((lambda x: setattr(x, "cell_size", 0.0), abi.AZ_ERR_INVALID),
                           (lambda x: setattr(x, "tstat_kind", abi.TSTAT_NOSE), abi.AZ_ERR_INVALID),   # Nose-Hoover without a relaxation time
                           (lambda x: setattr(x, "box", (0.0, 10.0, 10.0)), abi.AZ_ERR_INVALID)):
        t = azt.lj_system(ncell=(3, 3, 3))
        mutate(t)
        try:
            Engine(t, lib=lib)
        except AzError as err:
            assert err.status == status, (err.status, status)
        else:
            raise AssertionError("expected failure")
    t = azt.lj_system(ncell=(3, 3, 3))
    t.types = t.types.copy(); t.types[5] = 3
    try:
        Engine(t, lib=lib)
    except AzError as err:
        assert err.status == abi.AZ_ERR_INVALID
    else:
        raise AssertionError("expected failure")


ALL_CASES = [
    ("lj_basic", case_lj_basic),
    ("lj_anisotropic_grid", case_lj_anisotropic_even_odd_grid),
    ("lj_tiny_grids", case_lj_two_cells_per_axis_no_wrap),
    ("lj_wide_stencil", case_lj_small_cells_wide_stencil),
    ("lj_many_atoms_per_cell", case_lj_many_atoms_per_cell),
    ("ionic_buck_dsf", lambda lib: case_ionic(lib, abi.ELEC_FENNEL)),
    ("ionic_buck_direct", lambda lib: case_ionic(lib, abi.ELEC_DIRECT)),
    ("ionic_buck_ewald_real", lambda lib: case_ionic(lib, abi.ELEC_EWALD)),
    ("ionic_buck_nocoul", lambda lib: case_ionic(lib, abi.ELEC_NONE)),
    ("ionic_mixed_table_dsf", lambda lib: case_ionic(lib, abi.ELEC_FENNEL, kind="mixed")),
    ("ionic_elin_einv_bmh_ewald_real", lambda lib: case_ionic(lib, abi.ELEC_EWALD, kind="mixed2")),
    ("elin_einv_bmh_nocoul", lambda lib: case_ionic(lib, abi.ELEC_NONE, kind="mixed2")),
    ("ewald_reciprocal", case_ewald_reciprocal),
    ("ionic_efield_frozen", case_ionic_efield_frozen),
    ("surk_radiative_equil", case_surk_radiative),
    ("lj_radiative", case_lj_radiative),
    ("equilibration_scaling", case_equilibration_scaling),
    ("face_crossings_pressure", case_face_crossings_and_pressure),
    ("rebuild_interval_list", case_rebuild_interval_list),
    ("rebuild_interval_list_wide_stencils", case_rebuild_interval_list_wide_stencils),
    ("rebuild_interval_stale_tiles", case_rebuild_interval_stale_tiles),
    ("rebuild_interval_list_ionic", case_rebuild_interval_list_ionic),
    ("rebuild_interval_list_surk_radiative", case_rebuild_interval_list_surk_radiative),
    ("list_capacity", case_list_capacity_is_reported),
    ("skin_violation", case_skin_violation_is_reported),
    ("download_upload", case_download_upload_roundtrip),
    ("determinism", case_determinism),
    ("rdf_from_cells", case_rdf_from_cells_equals_brute_force),
    ("nose_hoover", case_nose_hoover),
    ("force_autocap", case_force_autocap_fires),
    ("bonds_angles", case_bonds_and_angles),
    ("bonds_reference_serial", case_bonds_match_the_reference_serial_build),
    ("reset_vels", case_reset_vels),
    ("tile_kernel_variant", case_tile_kernel_variant),
    ("list_every_step_variant", case_list_every_step_variant),
    ("fused_kick_drift", case_fused_kick_drift_is_bitwise_the_separate_kernels),
    ("tile_list_bitwise", case_tile_list_is_bitwise_the_global_rows),
    ("thread_kernel_variant", case_thread_kernel_variant),
    ("stencil_table_bitwise", case_stencil_table_is_bitwise_the_per_atom_stencil),
    ("graph_replay", case_graph_replay_is_bitwise_the_plain_launches),
    ("error_paths", case_error_paths),
]
