"""Domain-decomposition cases shared by the emulator tests (CPU, logic; tests/test_emu_bricks.py) and the
GPU tests (NCCL between real devices; tests/test_gpu_bricks.py).  N engines run as N host threads of one
process (ctypes releases the GIL; NCCL and the emulator mailbox both support one communicator rank per
thread); the result is compared, atom by atom through the persistent ids, with ONE engine holding the
whole system and with the oracle."""
import threading

import numpy as np
import pytest

import azt
from aztotmd_b200 import abi
from aztotmd_b200.engine import AzError, Engine, unique_comm_id


def run_bricks(lib, system, n_ranks, pgrid, nsteps, distributed=False, devices=None, rdf=None, one_call=False, **kw):
    """Advance `system` on n_ranks bricks; returns per-rank results gathered after the last step.  one_call: all steps in ONE
    az_engine_step call (the integrator then rides on the list kernel and the position streams rotate: the direct halo's route)."""
    if devices is not None and n_ranks > len(devices):
        # one rank per device over NCCL: with fewer devices the ranks that do exist would wait in ncclCommInitRank forever
        pytest.skip("needs %d GPUs, %d available" % (n_ranks, len(devices)))
    cid = unique_comm_id(lib)
    out = [None] * n_ranks
    err = [None] * n_ranks
    barrier = threading.Barrier(n_ranks)

    def work(r):
        try:
            s = system
            if distributed:  # every rank passes only the atoms of its own slab along the first decomposed axis, with ids
                s = slab_share(system, r, n_ranks, pgrid)
            e = Engine(s, lib=lib, rank=r, n_ranks=n_ranks, pgrid=pgrid, comm_id=cid,
                       device=(devices[r] if devices else 0), **kw)
            # all engines exist before any rank starts communicating: in ONE process a cudaMalloc on one
            # device can wait for a peer device whose NCCL kernel waits for this rank (NCCL's documented
            # multi-thread caveat); with one process per GPU (bench.py, production) this cannot happen
            barrier.wait()
            e.set_option("energy_every_step", 1)
            if rdf:
                e.rdf_begin(*rdf)
            stats = []
            if one_call:
                e.step(0, nsteps)
                stats = [None] * (nsteps - 1) + [e.stats()]
            for k in range(0 if one_call else nsteps):
                e.step(k, 1)
                stats.append(e.stats())
                if rdf:
                    e.rdf_sample()
            d = e.dump()
            a = e.download(abi.ORDER_SORTED)
            out[r] = dict(stats=stats, dump=d, atoms=a, info=e.info)
            if rdf:
                out[r]["rdf"] = e.rdf_counts()  # collective: the histogram of the whole box on every rank
            barrier.wait()
            e.close()
        except Exception as ex:  # noqa: BLE001
            err[r] = ex
            try:
                barrier.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(r,)) for r in range(n_ranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    real = [ex for ex in err if ex is not None and not isinstance(ex, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    for ex in err:
        if ex is not None:
            raise ex
    return out


def gather_by_id(results, n, key_src="dump"):
    frc = np.zeros((n, 3), dtype=np.float32)
    pos = np.zeros((n, 3), dtype=np.float64)
    vel = np.zeros((n, 3), dtype=np.float64)
    seen = np.zeros(n, dtype=np.int32)
    for r in results:
        ids = r["dump"]["ids"]
        frc[ids] = r["dump"]["frc"]
        seen[ids] += 1
        pos[r["atoms"]["ids"]] = r["atoms"]["pos"]
        vel[r["atoms"]["ids"]] = r["atoms"]["vel"]
    return frc, pos, vel, seen


def check_against_single(lib, system, n_ranks, pgrid, nsteps=3, tol=2e-5, distributed=False, devices=None, **kw):
    res = run_bricks(lib, system, n_ranks, pgrid, nsteps, distributed=distributed, devices=devices, **kw)
    e1 = Engine(system, lib=lib, **kw)
    e1.set_option("energy_every_step", 1)
    o = azt.Oracle(system)
    n = len(system.pos)
    for k in range(nsteps):
        e1.step(k, 1)
        o.step(k, 1)
        s1 = e1.stats()
        for r in res:  # every rank reports the same global scalars
            sr = r["stats"][k]
            sc = max(abs(s1.engVdW_d), abs(s1.engCoul1_d), abs(s1.engKin_d), 1e-9)
            assert abs(sr.engVdW_d - s1.engVdW_d) <= tol * sc, (k, sr.engVdW_d, s1.engVdW_d)
            assert abs(sr.engCoul1_d - s1.engCoul1_d) <= tol * sc
            assert abs(sr.engKin_d - s1.engKin_d) <= tol * sc
            assert abs(sr.engCoul2_d - s1.engCoul2_d) <= tol * sc and abs(sr.engCoul3 - s1.engCoul3) <= tol * max(sc, abs(s1.engCoul3))
            assert abs(sr.engTot - s1.engTot) <= 4 * tol * max(sc, abs(s1.engTot))
            assert sr.n_pairs == s1.n_pairs == o.pairs_in_range(), (sr.n_pairs, s1.n_pairs)
    frc, pos, vel, seen = gather_by_id(res, n)
    assert (seen == 1).all(), "every atom is owned by exactly one brick (%d missing, %d duplicated)" % ((seen == 0).sum(), (seen > 1).sum())
    d1 = azt.sort_by_id(e1.dump())
    do = azt.sort_by_id(o.dump())
    fs = np.zeros(n, dtype=np.float32)
    fs[o.dump()["ids"]] = o.force_scale()
    fs = np.maximum(fs, 1e-10)
    assert (np.abs(frc - d1["frc"]).max(1) <= 1e-4 * fs).all(), "brick forces differ from the single-GPU engine"
    assert (np.abs(frc - do["frc"]).max(1) <= 1e-4 * fs).all(), "brick forces differ from the oracle"
    a1 = e1.download(abi.ORDER_ORIGINAL)
    L = np.asarray(system.box)
    dp = np.abs(pos - a1["pos"])
    dp = np.minimum(dp, L - dp)
    assert dp.max() <= 1e-4
    assert np.abs(vel - a1["vel"]).max() <= 1e-4 * max(np.abs(a1["vel"]).max(), 1e-9)
    e1.close()
    return res


def case_two_bricks_along_x_match_single_engine(lib, devices=None):
    s = azt.lj_system(ncell=(8, 5, 5), seed=21)  # 5 x 3 x 3 cells of >= rc: bricks of 2 and 3 cells
    res = check_against_single(lib, s, 2, (2, 1, 1), nsteps=3, devices=devices)
    assert sum(len(r["dump"]["ids"]) for r in res) == len(s.pos)


def case_two_bricks_along_z_and_y(lib, devices=None):
    s = azt.lj_system(ncell=(5, 5, 8), seed=22)
    check_against_single(lib, s, 2, (1, 1, 2), nsteps=2, devices=devices)
    s = azt.lj_system(ncell=(5, 8, 5), seed=23)
    check_against_single(lib, s, 2, (1, 2, 1), nsteps=2, devices=devices)


def case_four_bricks_2x2(lib, devices=None):
    s = azt.lj_system(ncell=(8, 8, 5), seed=24)
    check_against_single(lib, s, 4, (2, 2, 1), nsteps=3, devices=devices)


def case_eight_bricks_2x2x2_ionic(lib, devices=None):
    s = azt.nacl_system(ncell=(7, 7, 7), elec=abi.ELEC_FENNEL, r_real=9.0, rc_vdw=9.0, cell_size=9.0)
    check_against_single(lib, s, 8, (2, 2, 2), nsteps=2, devices=devices)


def case_bricks_ewald_reciprocal(lib, devices=None):
    """'elec pme' on bricks: every rank sums q exp(ikr) over the atoms it owns, the structure factor is all-reduced,
    the k-space energy and the constant term are global on every rank."""
    s = azt.nacl_system(ncell=(7, 5, 5), elec=abi.ELEC_EWALD, alpha=0.3, jitter=0.4, ewald_k=(7, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.0)
    res = check_against_single(lib, s, 2, (2, 1, 1), nsteps=2, devices=devices)
    assert res[0]["stats"][0].engCoul2 > 1.0
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=2, distributed=True, devices=devices)


def case_bricks_rdf_histogram_is_global_and_exact(lib, devices=None):
    """SURVEY 8f-1 on bricks: every unordered pair is counted by the brick that owns the atom with the lower id; the
    all-reduced integer histogram equals the reference's O(N^2) brute_rdf counts (oracle) exactly."""
    for s, pg, r_max, dr in ((azt.lj_system(ncell=(8, 5, 5), seed=31), (2, 1, 1), 7.0, 0.1),
                             (azt.nacl_system(ncell=(7, 7, 5), elec=abi.ELEC_FENNEL, r_real=9.0, rc_vdw=9.0, cell_size=9.0), (2, 2, 1), 8.0, 0.05)):
        nr = pg[0] * pg[1] * pg[2]
        res = run_bricks(lib, s, nr, pg, 1, devices=devices, rdf=(r_max, dr))
        o = azt.Oracle(s)
        o.step(0, 1)
        counts, ns = res[0]["rdf"]
        ref = o.rdf_counts(r_max, dr, counts.shape[0], counts.shape[1])
        assert ns == 1 and counts.sum() > 1000
        for r in res:
            assert np.array_equal(r["rdf"][0], ref), (int(np.abs(r["rdf"][0].astype(np.int64) - ref.astype(np.int64)).sum()), int(ref.sum()))


def case_bricks_nose_hoover_uses_global_kinetic_energy(lib, devices=None):
    """SURVEY 8f-2 on bricks: the friction recurrence runs on the all-reduced kinetic energy, identically on every rank;
    energies, forces and velocities follow the single-GPU engine, with and without the equilibration rescale."""
    s = azt.lj_system(ncell=(8, 5, 5), seed=33, temperature=300.0, tstat=abi.TSTAT_NOSE)
    s.temperature = 120.0
    s.tstat_tau = 0.02
    res = check_against_single(lib, s, 2, (2, 1, 1), nsteps=5, devices=devices)
    ke0 = 1.5 * len(s.pos) * 300.0 * 8.617e-5
    assert res[0]["stats"][-1].engKin_d < 0.95 * ke0  # the thermostat acted
    s = azt.lj_system(ncell=(8, 5, 5), seed=34, temperature=150.0, tstat=abi.TSTAT_NOSE, n_equil=4, eq_freq=2)
    s.tstat_tau = 0.05
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=5, devices=devices)


def case_bricks_verlet_list_between_rebuilds(lib, devices=None):
    s = azt.lj_system(ncell=(8, 5, 5), seed=25, cell_size=9.3)
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=7, rebuild_every=4, skin=0.8, tol=1e-4, devices=devices)


def case_bricks_peer_memory_halo_equals_send_recv(lib, devices=None):
    """The per-step ghost refresh has two transports: stores into the neighbour's staging buffer over peer memory with arrival
    flags (NVLink / NVSwitch; az_k_ghost_put, consumed either by az_k_ghost_update_wait or -- fused -- by the ghost-touching
    blocks of the list pass themselves, straight from the staging; or -- direct -- stores straight into the ghost slots of the
    neighbour's sorted position stream, az_k_ghost_put_direct) and ncclSend / ncclRecv.  Same ghosts either way: the
    four runs agree bit for bit, and each agrees with the single engine.  (Several non-rebuild steps in a row exercise the
    alternation of the two staging halves; the Verlet-list schedule overlaps the refresh with the ghost-free blocks.)"""
    import os
    s = azt.lj_system(ncell=(8, 5, 5), seed=29, cell_size=2.5 * azt.AR_SIGMA + 0.8)
    runs = {}
    for tag, env in (("direct", {}), ("fused", {"AZTOT_NO_DIRECT_HALO": "1"}), ("peer", {"AZTOT_NO_FUSED_HALO": "1"}), ("sendrecv", {"AZTOT_NO_PEER_HALO": "1"})):
        os.environ.update(env)
        try:
            runs[tag] = check_against_single(lib, s, 2, (2, 1, 1), nsteps=7, rebuild_every=4, skin=0.8, tol=1e-4, devices=devices)
        finally:
            for k in env:
                os.environ.pop(k, None)
    assert all(r["info"].peer_halo == 0 for r in runs["sendrecv"])
    if devices is None:  # emulator: the ranks are threads of one process, peer memory always exists
        assert all(r["info"].peer_halo == 1 for r in runs["peer"]) and all(r["info"].peer_halo == 2 for r in runs["fused"])
        assert all(r["info"].peer_halo == 3 for r in runs["direct"])
    for tag in ("direct", "fused", "peer"):
        for a, b in zip(runs[tag], runs["sendrecv"]):
            assert np.array_equal(a["dump"]["ids"], b["dump"]["ids"]) and np.array_equal(a["atoms"]["pos"], b["atoms"]["pos"]), tag
            if tag == "fused" and devices is None:  # (emulator only: scalar row loop in the ghost-touching blocks, packed elsewhere)
                assert np.abs(a["dump"]["frc"] - b["dump"]["frc"]).max() <= 1e-5 * np.abs(b["dump"]["frc"]).max()
            else:
                assert np.array_equal(a["dump"]["frc"], b["dump"]["frc"]), tag
    # the same four ways with all steps in ONE call: the integrator rides on the list kernel, the position streams rotate, and
    # the direct halo really stores into the neighbours' streams (single-step calls drift in place and take the staged route)
    calls = {}
    for tag, env in (("direct", {}), ("fused", {"AZTOT_NO_DIRECT_HALO": "1"}), ("peer", {"AZTOT_NO_FUSED_HALO": "1"}), ("sendrecv", {"AZTOT_NO_PEER_HALO": "1"})):
        os.environ.update(env)
        try:
            calls[tag] = run_bricks(lib, s, 2, (2, 1, 1), 9, devices=devices, one_call=True, rebuild_every=4, skin=0.8)
        finally:
            for k in env:
                os.environ.pop(k, None)
    for tag in ("direct", "fused", "peer"):
        for a, b in zip(calls[tag], calls["sendrecv"]):
            assert np.array_equal(a["dump"]["ids"], b["dump"]["ids"]) and np.array_equal(a["atoms"]["pos"], b["atoms"]["pos"]), tag
            if tag == "fused" and devices is None:
                # emulator build only: the ghost-touching blocks of the fused form take the scalar row loop, everybody else the
                # packed one (two partial sums per component) -- equal to rounding there, bitwise on the GPU
                assert np.abs(a["dump"]["frc"] - b["dump"]["frc"]).max() <= 1e-5 * np.abs(b["dump"]["frc"]).max()
            else:
                assert np.array_equal(a["dump"]["frc"], b["dump"]["frc"]) and a["stats"][-1].engVdW_d == b["stats"][-1].engVdW_d, tag
    e1 = Engine(s, lib=lib, rebuild_every=4, skin=0.8)
    e1.step(0, 9)
    s1 = e1.stats()
    for r in calls["direct"]:
        assert r["stats"][-1].n_pairs == s1.n_pairs and abs(r["stats"][-1].engVdW_d - s1.engVdW_d) <= 1e-5 * abs(s1.engVdW_d)
    e1.close()
    print("peer_halo:", [r["info"].peer_halo for r in runs["direct"]], [r["info"].peer_halo for r in runs["fused"]], [r["info"].peer_halo for r in runs["peer"]])


def case_eight_bricks_list_direct_halo(lib, devices=None):
    """2 x 2 x 2 bricks on the Verlet-list schedule, all steps in one call: every rank has all 26 directions in use (each
    neighbour along an axis is the same rank twice), the direct halo stores into seven different ranks' position streams,
    migration and ghost re-selection happen at the rebuilds in between.  Against ONE engine holding the whole box."""
    s = azt.lj_system(ncell=(8, 8, 8), seed=91, cell_size=9.3, temperature=300.0)
    kw = dict(rebuild_every=4, skin=0.8)
    for n_ranks, pgrid in ((8, (2, 2, 2)), (4, (1, 2, 2))):
        res = run_bricks(lib, s, n_ranks, pgrid, 7, devices=devices, one_call=True, **kw)  # rebuilds at steps 0 and 4, direct puts at 1-3, 5, 6
        e1 = Engine(s, lib=lib, **kw)
        e1.step(0, 7)
        s1 = e1.stats()
        n = len(s.pos)
        frc, pos, vel, seen = gather_by_id(res, n)
        assert (seen == 1).all()
        for r in res:
            st = r["stats"][-1]
            assert st.n_pairs == s1.n_pairs, (st.n_pairs, s1.n_pairs)
            assert abs(st.engVdW_d - s1.engVdW_d) <= 2e-5 * abs(s1.engVdW_d) and abs(st.engKin_d - s1.engKin_d) <= 2e-5 * abs(s1.engKin_d)
            if devices is None:
                assert r["info"].peer_halo == 3
        d1 = azt.sort_by_id(e1.dump())
        assert np.abs(frc - d1["frc"]).max() <= 1e-4 * np.abs(d1["frc"]).max()
        a1 = e1.download(abi.ORDER_ORIGINAL)
        L = np.asarray(s.box)
        dp = np.abs(pos - a1["pos"])
        dp = np.minimum(dp, L - dp)
        assert dp.max() <= 1e-4
        e1.close()


def case_bricks_tile_lists(lib, devices=None):
    """list_format 2 (tile lists, az_tlist.cuh) on bricks: footprints that contain ghost atoms, the ghost-touching tiles waiting on
    the arrival flags of the direct halo -- against one engine (single-step calls) and against the global rows (one call)."""
    s = azt.lj_system(ncell=(8, 5, 5), seed=25, cell_size=9.3)
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=5, rebuild_every=4, skin=0.8, tol=1e-4, devices=devices, list_format=2)
    a = run_bricks(lib, s, 2, (2, 1, 1), 7, devices=devices, one_call=True, rebuild_every=4, skin=0.8, list_format=2)
    b = run_bricks(lib, s, 2, (2, 1, 1), 7, devices=devices, one_call=True, rebuild_every=4, skin=0.8, list_format=1)
    for x, y in zip(a, b):
        assert x["info"].tile_list == 1 and y["info"].tile_list == 0
        assert x["stats"][-1].n_pairs == y["stats"][-1].n_pairs and np.array_equal(x["dump"]["ids"], y["dump"]["ids"])
        if devices is None:  # (emulator: the global rows take the packed row loop, equal to rounding)
            assert np.abs(x["dump"]["frc"] - y["dump"]["frc"]).max() <= 1e-5 * np.abs(y["dump"]["frc"]).max()
        else:
            assert np.array_equal(x["dump"]["frc"], y["dump"]["frc"]) and np.array_equal(x["atoms"]["pos"], y["atoms"]["pos"])


def case_bricks_migration_hot_atoms(lib, devices=None):
    # hot and fast: atoms cross brick faces every few steps
    s = azt.lj_system(ncell=(8, 5, 5), seed=26, temperature=3000.0)
    s.timestep = 0.004
    res = run_bricks(lib, s, 2, (2, 1, 1), nsteps=12, devices=devices)
    n = len(s.pos)
    frc, pos, vel, seen = gather_by_id(res, n)
    assert (seen == 1).all()
    e1 = Engine(s, lib=lib)
    e1.step(0, 12)
    a1 = e1.download(abi.ORDER_ORIGINAL)
    L = np.asarray(s.box)
    dp = np.abs(pos - a1["pos"])
    dp = np.minimum(dp, L - dp)
    assert dp.max() <= 2e-3  # 12 steps of chaotic divergence from rounding-order differences
    st = res[0]["stats"][-1]
    s1 = e1.stats()
    assert sum(st.specAcBoxPos[0]) + sum(st.specAcBoxNeg[0]) == sum(s1.specAcBoxPos[0]) + sum(s1.specAcBoxNeg[0])
    e1.close()


def case_brick_layout_errors(lib, devices=None):
    s = azt.lj_system(ncell=(4, 4, 4), seed=27)  # 2 cells per axis: too fine for two bricks
    cid = unique_comm_id(lib)
    with pytest.raises(AzError) as ei:
        Engine(s, lib=lib, rank=0, n_ranks=2, pgrid=(2, 1, 1), comm_id=cid)
    assert ei.value.status == abi.AZ_ERR_UNSUPPORTED
    with pytest.raises(AzError) as ei:
        Engine(azt.lj_system(ncell=(8, 5, 5)), lib=lib, rank=0, n_ranks=2, pgrid=(3, 1, 1), comm_id=cid)
    assert ei.value.status == abi.AZ_ERR_INVALID


def case_bricks_equilibration_rescale_uses_global_kinetic_energy(lib, devices=None):
    # temp_scale on bricks: the scale factor comes from the kinetic energy of ALL bricks (one all-reduce of a double)
    s = azt.lj_system(ncell=(8, 5, 5), seed=71, n_equil=10, eq_freq=2, temperature=150.0)
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=5, devices=devices)


def case_bricks_ionic_list_with_rider(lib, devices=None):
    # two species, DSF Coulomb, Verlet list between rebuilds, several steps per call are exercised by run_bricks' single steps;
    # here one multi-step call per rank, so that the integrator rides on the list kernel across the halo refresh
    s = azt.nacl_system(ncell=(8, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.6)
    n = len(s.pos)
    cid = unique_comm_id(lib)
    out = [None, None]
    err = []
    bar = threading.Barrier(2)

    def work(r):
        try:
            e = Engine(s, lib=lib, rank=r, n_ranks=2, pgrid=(2, 1, 1), comm_id=cid, device=(devices[r] if devices else 0), rebuild_every=3, skin=0.6)
            bar.wait()
            e.step(0, 7)
            st = e.stats()
            out[r] = (e.dump(), e.download(abi.ORDER_SORTED), st)
            bar.wait()
            e.close()
        except Exception as ex:  # noqa: BLE001
            err.append(ex)
            bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    real = [x for x in err if not isinstance(x, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    assert not err
    e1 = Engine(s, lib=lib, rebuild_every=3, skin=0.6)
    e1.step(0, 7)
    s1 = e1.stats()
    frc = np.zeros((n, 3), np.float32)
    seen = np.zeros(n, np.int32)
    for d, a, st in out:
        frc[d["ids"]] = d["frc"]; seen[d["ids"]] += 1
        sc = max(abs(s1.engVdW_d), abs(s1.engCoul1_d), abs(s1.engKin_d))
        assert abs(st.engVdW_d - s1.engVdW_d) <= 1e-4 * sc and abs(st.engCoul1_d - s1.engCoul1_d) <= 1e-4 * sc and abs(st.engKin_d - s1.engKin_d) <= 1e-4 * sc
        assert st.n_pairs == s1.n_pairs
    assert (seen == 1).all()
    d1 = azt.sort_by_id(e1.dump())
    assert np.abs(frc - d1["frc"]).max() <= 2e-4 * np.abs(d1["frc"]).max()
    e1.close()


def slab_share(system, r, n_ranks, pgrid):
    """Distributed input: rank r passes the atoms whose coordinate along the first decomposed axis falls
    in its 1/P slab of the box (NOT the engine's cell-aligned brick: stragglers migrate in the first step)."""
    import copy
    ax = [a for a in range(3) if pgrid[a] > 1][0]
    P = pgrid[ax]
    pc = (r // (pgrid[1] * pgrid[2]), (r // pgrid[2]) % pgrid[1], r % pgrid[2])
    L = system.box[ax]
    x = np.asarray(system.pos)[:, ax] % L
    mine = (x >= L * pc[ax] / P) & (x < L * (pc[ax] + 1) / P)
    for a in range(3):  # other decomposed axes: exact halves as well
        if a != ax and pgrid[a] > 1:
            y = np.asarray(system.pos)[:, a] % system.box[a]
            mine &= (y >= system.box[a] * pc[a] / pgrid[a]) & (y < system.box[a] * (pc[a] + 1) / pgrid[a])
    s = copy.copy(system)
    idx = np.nonzero(mine)[0]
    s.pos = np.asarray(system.pos)[idx]
    s.types = np.asarray(system.types)[idx]
    s.vel = np.asarray(system.vel)[idx] if system.vel is not None else None
    s.global_ids = idx.astype(np.int32)
    s.n_atoms_global = len(system.pos)
    return s


def case_distributed_input_with_ids(lib, devices=None):
    s = azt.lj_system(ncell=(8, 5, 5), seed=28)
    check_against_single(lib, s, 2, (2, 1, 1), nsteps=2, distributed=True, devices=devices)
    s = azt.lj_system(ncell=(8, 8, 5), seed=29)
    check_against_single(lib, s, 4, (2, 2, 1), nsteps=2, distributed=True, devices=devices)


ALL_CASES = [(n[5:], f) for n, f in sorted(globals().items()) if n.startswith("case_") and callable(f)]
