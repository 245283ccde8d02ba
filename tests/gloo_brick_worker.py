"""Worker of tests/test_gloo_bricks.py: one process per rank (torchrun, backend gloo, CPU only).  Each rank
runs ONE engine of the emulator build; the halo exchange, the migration and the all-reduces of the C++ domain
decomposition travel over torch.distributed send/recv/all_reduce through the callback transport of
tests/emu/az_comm_emu.cpp.  Rank 0 also runs the whole system on one engine and compares atom by atom."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import azt  # noqa: E402
from aztotmd_b200 import abi  # noqa: E402
from aztotmd_b200.engine import Engine  # noqa: E402

EX_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))
AR_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int)


def as_tensor(ptr, nbytes):
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = azt.emu_lib()

    def exchange(n_ops, peers, send, send_bytes, recv, recv_bytes):
        try:
            reqs, nsend, nrecv = [], {}, {}
            for k in range(n_ops):  # the k-th send to a peer matches that peer's k-th receive from me: tag = ordinal
                p = peers[k]
                if send_bytes[k]:
                    t = nsend.get(p, 0); nsend[p] = t + 1
                    reqs.append(dist.isend(as_tensor(send[k], send_bytes[k]), dst=p, tag=t))
            for k in range(n_ops):
                p = peers[k]
                if recv_bytes[k]:
                    t = nrecv.get(p, 0); nrecv[p] = t + 1
                    reqs.append(dist.irecv(as_tensor(recv[k], recv_bytes[k]), src=p, tag=t))
            for r in reqs:
                r.wait()
            return 0
        except Exception as ex:  # noqa: BLE001
            print("exchange callback failed:", ex, flush=True)
            return 1

    def allreduce(data, n, take_max):
        try:
            t = torch.frombuffer((C.c_double * n).from_address(C.addressof(data.contents)), dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX if take_max else dist.ReduceOp.SUM)
            return 0
        except Exception as ex:  # noqa: BLE001
            print("allreduce callback failed:", ex, flush=True)
            return 1

    ex_cb, ar_cb = EX_CB(exchange), AR_CB(allreduce)
    lib.az_emu_set_comm_callbacks.argtypes = [EX_CB, AR_CB]
    lib.az_emu_set_comm_callbacks.restype = None
    lib.az_emu_set_comm_callbacks(ex_cb, ar_cb)

    cases = [("lj 2x1x1 every step", azt.lj_system(ncell=(8, 5, 5), seed=51), (2, 1, 1), {}, 3),
             ("lj 1x1x2 list", azt.lj_system(ncell=(5, 5, 8), seed=52, cell_size=9.3), (1, 1, 2), dict(rebuild_every=4, skin=0.8), 6),
             ("ionic 2x1x1", azt.nacl_system(ncell=(7, 5, 5), elec=abi.ELEC_FENNEL, r_real=9.0, rc_vdw=9.0, cell_size=9.0), (2, 1, 1), {}, 2)]
    # SURVEY 8f rows on bricks: the all-reduced structure factor of the Ewald sum, the Nose-Hoover friction on the
    # all-reduced kinetic energy, the RDF histogram all-reduced at read time
    pme = azt.nacl_system(ncell=(7, 5, 5), elec=abi.ELEC_EWALD, alpha=0.3, jitter=0.4, ewald_k=(7, 5, 5), r_real=9.0, rc_vdw=9.0, cell_size=9.0)
    nose = azt.lj_system(ncell=(8, 5, 5), seed=53, temperature=300.0, tstat=abi.TSTAT_NOSE)
    nose.temperature = 120.0
    nose.tstat_tau = 0.02
    cases += [("ionic pme 2x1x1", pme, (2, 1, 1), {}, 2), ("lj nose-hoover + rdf 2x1x1", nose, (2, 1, 1), {}, 3)]
    for name, s, pgrid, kw, nsteps in cases:
        rdf = (7.0, 0.1) if "rdf" in name else None
        e = Engine(s, lib=lib, rank=rank, n_ranks=world, pgrid=pgrid, comm_id=b"\0" * 128, **kw)
        e.set_option("energy_every_step", 1)
        if rdf:
            e.rdf_begin(*rdf)
        stats = []
        for k in range(nsteps):
            e.step(k, 1)
            stats.append(e.stats())
            if rdf and k == 0:
                e.rdf_sample()
        rdf_counts = e.rdf_counts()[0] if rdf else None  # collective
        d = e.dump()
        a = e.download(abi.ORDER_SORTED)
        mine = dict(ids=d["ids"], frc=d["frc"], pos=a["pos"], vel=a["vel"], aid=a["ids"])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        e.close()
        if rank == 0:
            lib.az_emu_set_comm_callbacks(EX_CB(), AR_CB())  # single engine below: no transport
            e1 = Engine(s, lib=lib, **kw)
            e1.set_option("energy_every_step", 1)
            if rdf:
                e1.rdf_begin(*rdf)
            n = len(s.pos)
            for k in range(nsteps):
                e1.step(k, 1)
                s1 = e1.stats()
                if rdf and k == 0:
                    e1.rdf_sample()
                    assert np.array_equal(rdf_counts, e1.rdf_counts()[0]) and rdf_counts.sum() > 1000, name
                assert abs(stats[k].engCoul2_d - s1.engCoul2_d) <= 1e-4 * max(abs(s1.engCoul2_d), 1e-9) and abs(stats[k].engCoul3 - s1.engCoul3) <= 1e-4 * max(abs(s1.engCoul3), 1e-9), name
                sc = max(abs(s1.engVdW_d), abs(s1.engCoul1_d), abs(s1.engKin_d), 1e-9)
                assert abs(stats[k].engVdW_d - s1.engVdW_d) <= 1e-4 * sc and abs(stats[k].engKin_d - s1.engKin_d) <= 1e-4 * sc, name
                assert stats[k].n_pairs == s1.n_pairs, (name, k, stats[k].n_pairs, s1.n_pairs)
            frc = np.zeros((n, 3), np.float32); seen = np.zeros(n, np.int32); pos = np.zeros((n, 3)); vel = np.zeros((n, 3))
            for g in gathered:
                frc[g["ids"]] = g["frc"]; seen[g["ids"]] += 1; pos[g["aid"]] = g["pos"]; vel[g["aid"]] = g["vel"]
            assert (seen == 1).all(), name
            d1 = azt.sort_by_id(e1.dump())
            assert np.abs(frc - d1["frc"]).max() <= 2e-4 * np.abs(d1["frc"]).max(), name
            a1 = e1.download(abi.ORDER_ORIGINAL)
            L = np.asarray(s.box)
            dp = np.abs(pos - a1["pos"]); dp = np.minimum(dp, L - dp)
            assert dp.max() <= 1e-4 and np.abs(vel - a1["vel"]).max() <= 1e-4 * np.abs(a1["vel"]).max(), name
            e1.close()
            lib.az_emu_set_comm_callbacks(ex_cb, ar_cb)
            print("ok:", name, flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
