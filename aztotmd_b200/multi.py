"""Multi-GPU driver of the benchmark: one process per GPU (torchrun), one engine per rank, bricks of
the periodic box with ghost-atom halo exchange over NCCL inside the CUDA library.

torch.distributed (gloo) is used only as the control plane -- to hand the 128-byte communicator id
of az_comm_unique_id to every rank, for the barriers around the timed region and for the max over
ranks of the device timings.  All data-path communication (migration at rebuild, ghost positions
every step, the reduction of the statistics) is NCCL send/recv/all-reduce issued by
libaztot_b200.so on the engine's stream.

Weak scaling (BASELINE config 5): every GPU holds a 100 x 100 x 100 fcc block
(4 000 000 argon atoms); 2 GPUs = 2x1x1 bricks, 4 = 2x2x1, 8 = 2x2x2 (32 M atoms).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import time

import numpy as np

BRICKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def brick_system(ncell, pgrid, rank, rebuild, skin, seed=12345):
    """This rank's share of the global fcc LJ argon lattice (pgrid * ncell unit cells per axis)."""
    from bench import AR_EPS, AR_MASS, AR_SIGMA
    from .engine import KB, M_SCALE, Species, System
    a = (4.0 / 0.8442) ** (1.0 / 3.0) * AR_SIGMA
    pc = (rank // (pgrid[1] * pgrid[2]), (rank // pgrid[2]) % pgrid[1], rank % pgrid[2])
    gcells = [ncell * pgrid[k] for k in range(3)]
    box = np.array([g * a for g in gcells])
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    ix, iy, iz = np.meshgrid(np.arange(ncell) + pc[0] * ncell, np.arange(ncell) + pc[1] * ncell, np.arange(ncell) + pc[2] * ncell, indexing="ij")
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    gid = (((cells[:, 0] * gcells[1] + cells[:, 1]) * gcells[2] + cells[:, 2])[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
    pos = (cells[:, None, :].astype(np.float64) + basis[None, :, :] + 0.25).reshape(-1, 3) * a
    rng = np.random.default_rng(1000 + rank)
    pos += rng.uniform(-0.05, 0.05, size=pos.shape)
    pos %= box
    n = len(pos)
    n_global = 4 * gcells[0] * gcells[1] * gcells[2]
    rc = 2.5 * AR_SIGMA
    s = System(pos=pos, types=np.zeros(n, dtype=np.int32), box=box, species=[Species("Ar", AR_MASS)], timestep=0.001,
               cell_size=rc + (skin if rebuild > 1 else 0.0), temperature=86.6, seed=seed,
               global_ids=gid.astype(np.int32), n_atoms_global=n_global)
    s.set_vdw(0, 0, "lnjs", rc, [AR_EPS, AR_SIGMA])
    m = AR_MASS * M_SCALE
    v = rng.standard_normal((n, 3)) * math.sqrt(KB * 86.6 / m)
    v -= v.mean(0)
    s.vel = v
    return s


def bench_main(args, config):
    import torch.distributed as dist
    from . import abi
    from .engine import Engine, unique_comm_id
    import bench

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must be launched with %d ranks (torchrun); WORLD_SIZE is %d" % (args.gpus, args.gpus, world))
    if world not in BRICKS:
        raise SystemExit("bench.py: supported GPU counts are 1, 2, 4, 8")
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
    import sys
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = abi.load_library()
    if lib.az_device_count() <= local:
        raise SystemExit("bench.py: rank %d has no CUDA device %d (the engine has no CPU fallback)" % (rank, local))

    # communicator id: rank 0 creates it, everyone gets the same 128 bytes
    box = [unique_comm_id(lib) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cid = box[0]

    pgrid = BRICKS[world]
    ncell, desc = bench.WORKLOADS[args.workload]
    s = brick_system(ncell, pgrid, rank, args.rebuild, args.skin)
    n_local = len(s.pos)
    n_global = s.n_atoms_global
    e = Engine(s, device=local, rebuild_every=args.rebuild, skin=args.skin if args.rebuild > 1 else 0.0,
               rank=rank, n_ranks=world, pgrid=pgrid, comm_id=cid)
    if getattr(args, "no_halo_overlap", False):
        e.set_option("halo_overlap", 0)
    K, W = args.steps, args.warmup
    # ---- halo_check: the bricks against ONE engine holding the same global lattice.  Two steps from the common start: step 0
    # bins, migrates and selects ghosts, step 1 runs on refreshed ghost positions (the per-step halo path); energies and the
    # pair count of step 1 must agree with the single-GPU run.
    HC_STEPS = 2
    e.set_option("energy_every_step", 1)
    e.step(0, HC_STEPS)
    st_b = e.stats()
    e.set_option("energy_every_step", 0)
    halo_check = None
    if rank == 0 and not getattr(args, "no_halo_check", False):
        try:
            parts = [s] + [brick_system(ncell, pgrid, r, args.rebuild, args.skin) for r in range(1, world)]
            order = np.argsort(np.concatenate([p.global_ids for p in parts]), kind="stable")
            g = parts[0]
            from .engine import System
            whole = System(pos=np.concatenate([p.pos for p in parts])[order], types=np.zeros(n_global, dtype=np.int32), box=g.box, species=g.species,
                           timestep=g.timestep, cell_size=g.cell_size, temperature=g.temperature, seed=g.seed)
            whole.vdw = g.vdw
            whole.vel = np.concatenate([p.vel for p in parts])[order]
            del parts
            one = Engine(whole, device=local, rebuild_every=args.rebuild, skin=args.skin if args.rebuild > 1 else 0.0)
            one.set_option("energy_every_step", 1)
            one.step(0, HC_STEPS)
            st_1 = one.stats()
            one.close()
            del whole
            rel = lambda a, b: abs(a - b) / max(abs(a), abs(b), 1e-30)  # noqa: E731
            d_v, d_k = rel(st_b.engVdW_d, st_1.engVdW_d), rel(st_b.engKin_d, st_1.engKin_d)
            pairs_ok = st_b.n_pairs == st_1.n_pairs  # unique in-range pairs: the all-reduced visits of the bricks / 2
            ok = d_v <= 2e-6 and d_k <= 2e-6 and pairs_ok
            halo_check = {"status": "ok" if ok else "MISMATCH", "steps": HC_STEPS, "engVdW_rel": d_v, "engKin_rel": d_k, "tolerance": 2e-6,
                          "pairs_bricks": int(st_b.n_pairs), "pairs_single_gpu": int(st_1.n_pairs),
                          "what": "%d bricks vs one engine holding the same %d-atom lattice on GPU 0" % (world, n_global)}
        except Exception as ex:  # noqa: BLE001 -- the check must not take the benchmark line down
            halo_check = {"status": "error", "why": repr(ex)[:300]}
    dist.barrier()
    e.step(HC_STEPS, W)
    st = e.stats()  # collective; surfaces overflow flags early
    launches0 = e.launch_count
    clocks = bench.ClockSampler(device=local) if rank == 0 else None
    if clocks:
        clocks.start()
    e.profile(2)
    dist.barrier()
    e.sync()
    ms = e.step_timed(HC_STEPS + W, K)  # CUDA events on the engine's stream, this rank
    e.sync()
    dist.barrier()
    tm = e.timers()
    e.profile(0)
    clk = clocks.stop() if clocks else None
    launches = e.launch_count - launches0
    st = e.stats()
    import torch
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0])
    value = n_global * K / (ms_max * 1e-3)

    # ---- end to end: every rank uploads its share from pinned host memory, steps once, reads the statistics
    n = n_local
    hp = lib.az_host_alloc(6 * 8 * n)
    host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(6, n))
    host[0:3] = np.asarray(s.pos).T
    host[3:6] = np.asarray(s.vel).T
    dp = C.POINTER(C.c_double)
    ptr = [C.cast(hp + 8 * n * k, dp) for k in range(6)]
    ne = max(1, args.e2e_steps)
    first = HC_STEPS + W + K
    up = lib.az_engine_upload_async
    for _ in range(2):
        assert up(e._h, *ptr) == 0, lib.az_last_error(e._h)
        e.step(first, 1); e.stats(); first += 1
    e.sync()
    dist.barrier()
    t0 = time.perf_counter()
    assert up(e._h, *ptr) == 0
    for k in range(ne):
        e.step(first, 1)
        if k + 1 < ne:
            assert up(e._h, *ptr) == 0
        e.stats()
        first += 1
    e.sync()
    dist.barrier()
    t1 = time.perf_counter()
    t = torch.tensor([t1 - t0], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global * ne / float(t[0])
    lib.az_host_free(hp)

    if rank == 0:
        hbm_peak, hbm_src = bench.load_peaks()
        dom = "pair_list" if tm["pair_list"]["launches"] > 0 else "pair"
        # per step, not per launch: with the halo overlap the list pass is two launches (blocks without / with ghosts)
        pair_ms = tm[dom]["ms"] / K if dom == "pair_list" else tm[dom]["ms"] / max(1, tm[dom]["launches"])
        # rank 0's pair kernel against the FP32 peak: its share of the unique pairs is 1 / world
        tf, mhz = C.c_double(), C.c_double()
        lib.az_measure_fp32_peak(local, C.byref(tf), C.byref(mhz))
        flops = bench.FLOP_PER_PAIR * st.n_pairs / world + bench.FLOP_PER_ATOM_INTEGRATE * n_local  # the integrator rides on the list kernel
        achieved_tf = flops / (pair_ms * 1e-3) / 1e12
        config = dict(config)
        config.update({"workload": "%s per GPU x %d GPUs (bricks %dx%dx%d, %d atoms); halo exchange over NCCL; %s" % (
            desc, world, pgrid[0], pgrid[1], pgrid[2], n_global, config["workload"].split(";", 1)[1].strip()),
            "n_atoms": n_global, "bricks": list(pgrid)})
        out = {"metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config, "clocks": clk,
               "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": 6 * 8 * n * world,
                       "d2h_bytes_per_step": C.sizeof(abi.az_stats) * world, "steps": ne, "timer": "host wall clock, max over ranks"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "fp32", "kernel": "az_k_pair_list" if dom == "pair_list" else "az_k_pair", "achieved": achieved_tf,
                            "peak": tf.value, "unit": "TFLOP/s", "frac": achieved_tf / tf.value if tf.value else None, "traffic": None,
                            "peak_source": "az_measure_fp32_peak on rank 0 in this run", "avg_launch_ms": pair_ms, "launches": tm[dom]["launches"],
                            "algorithmic": "%.0f flop per unique in-cutoff pair x %d pairs / %d ranks + %.0f flop per atom of the integrator that rides on the kernel" % (
                                bench.FLOP_PER_PAIR, st.n_pairs, world, bench.FLOP_PER_ATOM_INTEGRATE),
                            "share_of_step": tm[dom]["ms"] / ms},
               "kernel_ms_per_step": {k: v["ms"] / K for k, v in tm.items()},
               "energy": {"engKin": st.engKin_d, "engVdW": st.engVdW_d, "pairs_in_cutoff": st.n_pairs},
               "halo_check": halo_check["status"] if halo_check else "skipped", "halo_check_detail": halo_check,
               "ns_per_day": 86400.0 / (ms_max / K * 1e-3) * 1e-6,
               "hbm_peak": {"value": hbm_peak, "source": hbm_src}}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    e.close()
    dist.destroy_process_group()
