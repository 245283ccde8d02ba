#!/usr/bin/env python
"""bench.py -- atom-steps/s of the MD hot path (bin+reorder -> pair forces -> Verlet+thermostat).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one MD time step (one pass of the whole hot path) over the workload's atoms.
N = 1 workload: BASELINE.json configs[2], synthetic fcc LJ argon, 4 000 000 atoms, rc = 2.5 sigma,
cell-list rebuild every 10 steps (seeded synthetic lattice + Maxwell velocities; no dataset exists).
`value`   : atoms * K / device time of K steps (CUDA events on the engine's stream), state resident in HBM.
`e2e`     : the same metric through the C ABI with HOST buffers: every step uploads positions and
            velocities from pinned host memory (az_engine_upload_async: the copy of step k+1's inputs
            overlaps the kernels of step k), advances one step and reads the step's statistics back
            (az_engine_stats); host wall clock around the loop.
`roofline`: the dominant kernel (pair forces) against the FP32 FMA peak measured in the same run
            (az_measure_fp32_peak; MEASURED_PEAKS.json has no FP32 figure), plus `roofline_hbm`
            for the bandwidth-bound remainder (integrate / sort) against the measured HBM copy peak.
`cpu_baseline` / --impl reference: the reference's own serial twin (oracle/_ref/ref_serial_harness,
            compiled from the reference's sources) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 30.0          # flops per unique in-cutoff LJ pair (both partners' updates)
FLOP_PER_ATOM_INTEGRATE = 40.0
HBM_BYTES_PER_ATOM_STEP = 112.0   # pos/frc through the pair kernel + two integrate passes
HBM_BYTES_PER_ATOM_SORT = 72.0    # per rebuild

WORKLOADS = {
    # name: (fcc cells per edge, description)
    "lj4m": (100, "synthetic fcc LJ argon, 4,000,000 atoms, rho*=0.8442, rc=2.5 sigma"),
    "lj500k": (50, "synthetic fcc LJ argon, 500,000 atoms, rho*=0.8442, rc=2.5 sigma"),
    "lj4k": (10, "synthetic fcc LJ argon, 4,000 atoms, rho*=0.8442, rc=2.5 sigma"),
    "lj32k": (20, "synthetic fcc LJ argon, 32,000 atoms, rho*=0.8442, rc=2.5 sigma"),
    "lj131k": (32, "synthetic fcc LJ argon, 131,072 atoms, rho*=0.8442, rc=2.5 sigma (the deck of profiles/ref_cuda_probe.py)"),
    # BASELINE config 3 of the ionic kind: rock-salt NaCl-like melt, Buckingham + real-space Coulomb (DSF), 1 000 000 ions
    "nacl1m": (-50, "synthetic NaCl-like ionic melt, 1,000,000 ions (50^3 rock-salt cells, a = 5.9 A), Buckingham rc 9 A + DSF Coulomb rc 12 A alpha 0.2, 1200 K"),
    # the full Ewald sum ('elec pme 12 0.2 8 8 8': real space + 8x15x15 half space of k-vectors inside the cut-off + constant)
    "nacl64k_pme": (-20.5, "synthetic NaCl-like ionic melt, 64,000 ions (20^3 rock-salt cells), Buckingham rc 9 A + Ewald sum: real space rc 12 A alpha 0.2, reciprocal space kx=ky=kz=8"),
    "nacl1m_pme": (-50.5, "synthetic NaCl-like ionic melt, 1,000,000 ions, Buckingham rc 9 A + Ewald sum: real space rc 12 A alpha 0.2, reciprocal space kx=ky=kz=8 (the reference's k-vector budget, not a converged sum at this box size)"),
    # latency-bound decks of the reference (BASELINE configs 0 and 1), rebuilt synthetically (the shipped files cannot travel)
    "cs1": (0, "case-study-1-like: 40,000 argon atoms, dilute gas in a 1141.5 A box, LJ rc 4.0 A, cell_list 85, radiative thermostat 298 K"),
    "cs2": (0, "case-study-2-like: 4,000 abstract atoms, 35 A box, surk rc 6.0 A + radii, cell_list 2.7, radiative thermostat 500 K"),
}
AR_EPS, AR_SIGMA, AR_MASS = 0.01006, 3.3952, 39.9


def lj_system(ncell, rebuild, skin, seed=12345, cell_div=1):
    from aztotmd_b200.decks import fcc_positions
    from aztotmd_b200.engine import Species, System
    a = (4.0 / 0.8442) ** (1.0 / 3.0) * AR_SIGMA
    pos, box = fcc_positions((ncell,) * 3, a, jitter=0.05, seed=1)
    rc = 2.5 * AR_SIGMA
    s = System(pos=pos, types=np.zeros(len(pos), dtype=np.int32), box=box, species=[Species("Ar", AR_MASS)],
               timestep=0.001, cell_size=(rc + (skin if rebuild > 1 else 0.0)) / cell_div, temperature=86.6, seed=seed)
    s.set_vdw(0, 0, "lnjs", rc, [AR_EPS, AR_SIGMA])
    s.vel = s.gauss_velocities(seed)
    return s


def ionic_system(ncell, rebuild, skin, seed=2):
    from aztotmd_b200 import abi
    pme = ncell != int(ncell)  # x.5: the same box with the full Ewald sum
    ncell = int(ncell)
    from aztotmd_b200.decks import rocksalt_positions
    from aztotmd_b200.engine import Species, System
    pos, types, box = rocksalt_positions((ncell,) * 3, 5.9, jitter=0.05, seed=seed)
    s = System(pos=pos, types=types.astype(np.int32), box=box, species=[Species("Na", 22.99, 1.0), Species("Cl", 35.45, -1.0)],
               timestep=0.001, cell_size=12.0 + (skin if rebuild > 1 else 0.0), elec_kind=abi.ELEC_EWALD if pme else abi.ELEC_FENNEL, r_real=12.0, alpha=0.2,
               temperature=1200.0, seed=seed, ewald_k=(8, 8, 8) if pme else (0, 0, 0))
    s.set_vdw(0, 0, "buck", 9.0, [424.097, 0.317, 1.05])
    s.set_vdw(0, 1, "buck", 9.0, [1256.31, 0.317, 7.00])
    s.set_vdw(1, 1, "buck", 9.0, [3488.998, 0.317, 72.40])
    s.vel = s.gauss_velocities(seed)
    return s


def case_study_system(name, seed=12345):
    """Synthetic twins of the reference's two shipped decks (same sizes, box, potentials, thermostat, cell_list)."""
    from aztotmd_b200 import abi
    from aztotmd_b200.engine import Species, System
    rng = np.random.default_rng(seed)
    if name == "cs1":
        m, L = 35, 1141.5  # 35^3 = 42 875 sites of a simple cubic lattice, the first 40 000 are used
        g = (np.arange(m) + 0.5) * (L / m)
        pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:40000] + rng.uniform(-10.0, 10.0, size=(40000, 3))
        s = System(pos=pos % L, types=np.zeros(40000, dtype=np.int32), box=(L, L, L), species=[Species("Ar", AR_MASS, 0.0, 0, 2.73, 4.731, 0.2)],
                   timestep=0.001, cell_size=85.0, temperature=298.0, seed=seed, tstat_kind=abi.TSTAT_RADI)
        s.set_vdw(0, 0, "lnjs", 4.0, [AR_EPS, AR_SIGMA])
        return s
    g = np.arange(16) * 1.3
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:4000]
    pos = pos + (35.0 - g[-1]) / 2 + rng.uniform(-0.2, 0.2, size=pos.shape)
    s = System(pos=pos, types=np.zeros(4000, dtype=np.int32), box=(35.0, 35.0, 35.0), species=[Species("Ar", AR_MASS, 0.0, 0, 2.73, 4.731, 0.2)],
               timestep=0.001, cell_size=2.7, temperature=500.0, seed=seed, tstat_kind=abi.TSTAT_RADI, n_equil=10000, eq_freq=2500)
    s.set_vdw(0, 0, "surk", 6.0, [75.0, 8.0, 1.0, 1.0])
    return s


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_entry(kernel, args):
    """The committed `ncu --set full` capture of the dominant kernel on the same workload (profiles/r2/ncu_traffic.json, else
    profiles/r1/): dram__bytes_read.sum + dram__bytes_write.sum per launch and the pipe that limits it; None if not captured."""
    for rnd in ("r2", "r1"):
        p = os.path.join(ROOT, "profiles", rnd, "ncu_traffic.json")
        if not os.path.exists(p):
            continue
        with open(p) as f:
            d = json.load(f)
        key = "%s/rebuild%d/skin%.2f" % (args.workload, args.rebuild, args.skin if args.rebuild > 1 else 0.0)
        ent = d.get(key, {}).get(kernel)
        if ent:
            return dict(ent, source="profiles/%s/ncu_traffic.json" % rnd)
    return None


def ncu_traffic(kernel, args):
    ent = ncu_entry(kernel, args)
    return ent["dram_bytes_per_launch"] if ent else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def small_deck_line(lib, name="cs1", steps=2000, warmup=200):
    """The other half of BASELINE.json's metric ("LJ Argon at 40k and 4M atoms"): the reference's SHIPPED case study 1
    (tests/golden/decks/cs1: 40 000 argon atoms, dilute gas, LJ rc 4 A, radiative thermostat, cell_list 85) on its own
    schedule (bin + sort every step), as plain launches and as a CUDA-graph replay.  Latency-bound: one wave of CTAs."""
    from aztotmd_b200 import abi
    from aztotmd_b200.engine import Engine
    d = os.path.join(ROOT, "tests", "golden", "decks", name)
    if not os.path.isdir(d):
        return {"unavailable": "tests/golden/decks/%s missing" % name}
    h = C.c_void_p()
    if lib.az_deck_load(d.encode(), 12345, C.byref(h)) != 0:
        return {"unavailable": "deck does not parse"}
    desc = lib.az_deck_desc(h).contents
    what = {"cs1": "shipped case study 1 (tests/golden/decks/cs1): 40,000 Ar, LJ rc 4 A, radiative thermostat 298 K, rebuild every step",
            "cs2": "shipped case study 2 (tests/golden/decks/cs2): 4,000 atoms, radius/temperature-dependent surk potential + radiative thermostat 500 K, rebuild every step"}
    out = {"workload": what.get(name, name), "n_atoms": int(desc.n_atoms), "steps": steps}
    for tag, graph in (("plain", 0), ("graph", 1)):
        opts = abi.az_cuda_opts.from_buffer_copy(bytes(lib.az_deck_opts(h).contents))
        opts.use_graph = graph
        e = Engine(desc, lib=lib, opts=opts)
        e.step(0, warmup)
        e.stats()
        first = warmup
        ms = e.step_timed(first, steps); first += steps
        if graph:  # the first call of a shape records the graph; time a replay
            ms = min(e.step_timed(first + k * steps, steps) for k in range(2))
        out[tag] = {"us_per_step": ms / steps * 1e3, "atom_steps_per_s": desc.n_atoms * steps / (ms * 1e-3)}
        e.close()
    lib.az_deck_free(h)
    return out


def ionic_line(steps=40, warmup=10):
    """BASELINE config 4: synthetic NaCl-like ionic melt, 1 M ions, Buckingham + real-space (DSF) Coulomb, list rebuilt every 5 steps."""
    from aztotmd_b200.engine import Engine
    ncell, desc = WORKLOADS["nacl1m"]
    s = ionic_system(-ncell, 5, 0.6)  # (ionic workloads carry their cell count negated)
    e = Engine(s, rebuild_every=5, skin=0.6)
    e.step(0, warmup)
    e.stats()
    ms = e.step_timed(warmup, steps)
    st = e.stats()
    n = len(s.pos)
    out = {"workload": desc + "; rebuild every 5 steps, skin 0.6 A", "n_atoms": n, "steps": steps, "ms_per_step": ms / steps,
           "atom_steps_per_s": n * steps / (ms * 1e-3), "pairs_in_cutoff": int(st.n_pairs)}
    e.close()
    return out


def ref_cuda_baseline():
    """north_star's first target is a ratio to the reference's own CUDA build.  That build (oracle/_ref/aztotmd_ref, the
    reference sources compiled for sm_100a) was run on a B200 by profiles/ref_cuda_probe.py; the committed record
    says what it did there -- see profiles/r2/reference_cuda.json and DESIGN.md section 6."""
    p = os.path.join(ROOT, "profiles", "r2", "reference_cuda.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        j = json.load(f)
    t = j.get("timing", {})
    a = t.get("reference_cuda", {})
    return {"value": a.get("atom_steps_per_s"), "unit": "atom-steps/s", "n_atoms": t.get("n_atoms"),
            "stat_dat_valid": bool(a.get("stat", {}).get("valid")), "loop_completed": a.get("loop_completed"),
            "new_engine_same_deck": t.get("new_engine", {}).get("atom_steps_per_s"),
            "source": "profiles/r2/reference_cuda.json (gpurun job on a B200 of this pool, not this run)", "note": j.get("note")}


# ----------------------------------------------------------------------------- reference arm
def sample_cells_for(seconds, steps, rate=2.7e5):
    n_target = max(4000.0, min(256000.0, seconds * rate / max(1, steps)))
    m = int(round((n_target / 4.0) ** (1.0 / 3.0)))
    return max(12, m)  # >= 12 fcc cells: >= 8 linked cells per axis for the serial twin


def run_reference_sample(steps, warmup, budget_s=25.0):
    """Time the reference's serial twin (1 thread: the code is serial) on a bounded fcc LJ sample of
    the benchmark workload (same density, cutoff, potential; fewer atoms)."""
    from aztotmd_b200 import decks
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_serial_harness")
    m = sample_cells_for(budget_s, steps + warmup)
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory() as tmp:
            d = os.path.join(tmp, "deck")
            info = decks.lj_argon_deck(d, m, nstep=steps, init_vel="gaus")
            if warmup > 0:
                subprocess.run([harness, str(warmup), "fast", "-", "1"], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True)
            r = subprocess.run([harness, str(steps), "fast", "-", "1"], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                j = json.loads(line[-1])
                return dict(value=j["atom_steps_per_s"], ms_per_step=j["s_per_step"] * 1e3, kind="reference", cores=1,
                            n_atoms=info["n_atoms"],
                            sample="%d-atom fcc LJ argon cut of the workload (same density/cutoff), %d steps, "
                                   "reference serial twin (aztot_serial sources, g++ -O2), 1 thread" % (info["n_atoms"], steps))
    # the reference binary is absent: time the C restatement (oracle port) instead
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import azt
    s = azt.lj_system(ncell=(m, m, m))
    o = azt.Oracle(s)
    o.step(0, max(1, warmup))
    t0 = time.perf_counter()
    o.step(warmup, steps)
    dt = time.perf_counter() - t0
    return dict(value=len(s.pos) * steps / dt, ms_per_step=dt / steps * 1e3, kind="port", cores=1, n_atoms=len(s.pos),
                sample="%d-atom fcc LJ argon cut of the workload, %d steps, oracle/az_oracle.c, 1 thread" % (len(s.pos), steps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lj4m", choices=sorted(WORKLOADS))
    ap.add_argument("--rebuild", type=int, default=10)
    ap.add_argument("--skin", type=float, default=0.25)
    ap.add_argument("--pair-variant", type=int, default=0, help="0 auto (Verlet list between rebuilds), 1 cell tiles only")
    ap.add_argument("--graph", action="store_true", help="az_cuda_opts.use_graph: replay the timed call as one CUDA graph (per-kernel timers are then unavailable)")
    ap.add_argument("--option", action="append", default=[], help="engine option key=value (az_engine_set_option), for A/B runs")
    ap.add_argument("--list-dbg", type=int, default=0, help="diagnostic variants of the list pass (results are NOT physical)")
    ap.add_argument("--no-halo-overlap", action="store_true", help="multi-GPU: refresh the ghosts on the engine's stream (A/B of the overlap)")
    ap.add_argument("--cell-div", type=int, default=1, help="cell_list = (rc + skin) / cell_div")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-halo-check", action="store_true", help="multi-GPU: skip the comparison with a single-GPU run of the same lattice")
    ap.add_argument("--no-small-deck", action="store_true", help="skip the 40 000-atom case-study-1 leg of the default run")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncell, desc = WORKLOADS[args.workload]
    n_atoms = 4 * ncell ** 3 if ncell > 0 else (8 * int(-ncell) ** 3 if ncell < 0 else (40000 if args.workload == "cs1" else 4000))
    config = {"workload": "%s; cell_list = rc%s; rebuild every %d steps; dt = 1 fs; NVE" % (
        desc, " + skin %.2f A" % args.skin if args.rebuild > 1 else "", args.rebuild),
        "n_atoms": n_atoms, "rebuild_every": args.rebuild, "skin": args.skin if args.rebuild > 1 else 0.0,
        "l2": ("inputs larger than L2 (%.0f MB of pos/vel/frc streams vs 126 MB)" % (n_atoms * 48 / 1e6)) if n_atoms * 48 > 126e6 else
              "latency-bound deck: the whole state (%.1f MB) fits L2 by construction of the reference's case study; no flush" % (n_atoms * 48 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference_sample(args.steps, args.warmup)
        # the serial twin's start-up is O(N^2) (all_pairs, src/sys_init.cpp:1184): it is timed on a bounded cut of the workload
        # (same lattice, density, potential, cut-off) and the line says so -- n_atoms is the cut's, not the 4 M of the GPU arm
        config = dict(config, n_atoms=int(r["n_atoms"]), n_atoms_gpu_arm=n_atoms,
                      workload=config["workload"] + "; CPU arm: %d-atom cut of this lattice" % r["n_atoms"])
        print(json.dumps({"impl": "reference", "metric": "atom-steps/s", "value": r["value"], "unit": "atom-steps/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config, "ns_per_day": 86400.0 / (r["ms_per_step"] * 1e-3) * 1e-6,
                          "cpu_baseline": {"value": r["value"], "unit": "atom-steps/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                          "e2e": {"value": r["value"], "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if world > 1 or args.gpus > 1:
        from aztotmd_b200 import multi
        multi.bench_main(args, config)
        return

    from aztotmd_b200 import abi
    from aztotmd_b200.engine import Engine
    lib = abi.load_library()
    if lib.az_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    hbm_peak, hbm_src = load_peaks()
    tf, mhz = C.c_double(), C.c_double()
    assert lib.az_measure_fp32_peak(0, C.byref(tf), C.byref(mhz)) == 0
    fp32_peak = tf.value

    if ncell > 0:
        s = lj_system(ncell, args.rebuild, args.skin, cell_div=args.cell_div)
    elif ncell < 0:
        s = ionic_system(-ncell, args.rebuild, args.skin)
    else:
        s = case_study_system(args.workload)
    e = Engine(s, rebuild_every=args.rebuild, skin=args.skin if args.rebuild > 1 else 0.0, pair_variant=args.pair_variant, use_graph=args.graph)
    if args.list_dbg:
        e.set_option("list_dbg", args.list_dbg)
    for kv in args.option:
        k, v = kv.split("=")
        e.set_option(k, float(v))
    K, W = args.steps, args.warmup
    e.step(0, W)
    st = e.stats()  # also surfaces skin/overflow errors early
    pairs = st.n_pairs
    launches0 = e.launch_count
    clocks = ClockSampler()
    clocks.start()
    # the timed region: exactly K steps after the W warm-up steps.  Per-section CUDA events (kernel_ms_per_step, the roofline
    # of the dominant kernel) are recorded INSIDE it for the large workloads, where they cost nothing measurable; on the
    # latency-bound decks (and for a graph replay) they would add launch gaps of their own, so those get a second pass.
    sections_inside = (not args.graph) and n_atoms >= 200000
    launches0 = e.launch_count
    first = W
    if sections_inside:
        e.profile(2)
    ms = e.step_timed(first, K)
    first += K
    if sections_inside:
        tm = e.timers()
        e.profile(0)
        ms_prof = ms
    launches = e.launch_count - launches0
    graph = None
    if args.graph:
        # az_cuda_opts.use_graph: the call above captured + instantiated the graph of this call shape; time a replay
        info = abi.az_engine_info_t()
        seen = 0
        for _ in range(12):  # the buffer rotation of the list path returns to the same arrangement after a few calls
            launches0 = e.launch_count
            ms = e.step_timed(first, K)
            first += K
            launches = e.launch_count - launches0
            lib.az_engine_info(e._h, C.byref(info))
            if info.graph_replays > seen:
                break
            seen = info.graph_replays
        assert info.graph_replays >= 1, "no call of this shape was replayed from a captured graph"
        graph = {"replays": int(info.graph_replays), "graphs_cached": int(info.graphs_cached)}
    clk = clocks.stop()
    st = e.stats()
    value = n_atoms * K / (ms * 1e-3)
    if not sections_inside:  # second pass of K steps with the section events
        e.profile(2)
        ms_prof = e.step_timed(first, K)
        first += K
        tm = e.timers()
        e.profile(0)

    # dominant kernel: the Verlet-list pass when it runs (9 of 10 steps), else the cell-tile sweep
    dom = "pair_list" if tm["pair_list"]["launches"] > 0 else "pair"
    info = e.refresh_info()
    tile_list = bool(getattr(info, "tile_list", 0))
    dom_kernel = ("az_k_pair_tlist" if tile_list else "az_k_pair_list") if dom == "pair_list" else "az_k_pair"
    pair_ms = tm[dom]["ms"] / max(1, tm[dom]["launches"])
    pair_total_ms = tm["pair"]["ms"] + tm["pair_list"]["ms"]
    rest_ms = (tm["integrate_a"]["ms"] + tm["integrate_b"]["ms"] + tm["scan"]["ms"] + tm["sort"]["ms"]) / K
    # algorithmic flop per unique in-range pair: LJ 30; Buckingham + DSF 66 (each transcendental counted as 8);
    # surk 44 (30 - 9 for LJ's polynomial + sqrt and 4 divisions at ~5 each + 3)
    flop_per_pair = {"nacl1m": 66.0, "nacl1m_pme": 66.0, "nacl64k_pme": 66.0, "cs2": 44.0}.get(args.workload, FLOP_PER_PAIR)
    flops_per_launch = flop_per_pair * st.n_pairs
    rides = dom == "pair_list" and args.workload not in ("cs1", "cs2")  # thermostat-free: the integrator rides on the list kernel
    if rides:
        flops_per_launch += FLOP_PER_ATOM_INTEGRATE * n_atoms
    achieved_tf = flops_per_launch / (pair_ms * 1e-3) / 1e12
    # HBM-bound remainder, graded on the bytes its kernels really move.  With the rider the integrate passes are gone (the
    # list kernel reads / writes pos and vel itself) except az_k_integrate_a at the head and az_k_integrate_b at the tail of
    # a call, so what is left per step is the reorder at a rebuild: az_k_sort_keys + az_k_sort_scatter, 72 B per atom
    # ((16 + 16 + 4) B read and written).  Without the rider az_k_integrate_ba adds 80 B per atom and step.
    n_sorts = tm["sort"]["launches"] // 2
    sort_ms = tm["sort"]["ms"] / max(1, n_sorts)
    sort_gbs = HBM_BYTES_PER_ATOM_SORT * n_atoms / (sort_ms * 1e-3) / 1e9 if n_sorts else 0.0
    hbm_kernels = "az_k_sort_keys + az_k_sort_scatter (per rebuild)"
    achieved_gbs, hbm_ms, hbm_bytes = sort_gbs, sort_ms, HBM_BYTES_PER_ATOM_SORT * n_atoms
    if not rides and tm["integrate_b"]["launches"] > 0:  # the fused second-kick + first-kick pass of every step (incl. az_k_finalize)
        ib_ms = tm["integrate_b"]["ms"] / K
        hbm_kernels = "az_k_integrate_ba (+ az_k_finalize) per step"
        achieved_gbs, hbm_ms, hbm_bytes = 80.0 * n_atoms / (ib_ms * 1e-3) / 1e9, ib_ms, 80.0 * n_atoms

    # ---- end to end through the C ABI with HOST buffers (pinned): every step uploads this step's positions and velocities,
    # advances one step and reads the step's result back.  Three variants of the same loop:
    #   f64      the reference's own host layout (struct Atoms: six double arrays, 48 B per atom) in, az_stats out
    #   f32      float arrays in (az_engine_upload_f32_async, 24 B per atom), az_stats out            <- the `e2e` headline
    #   f32+pos  float arrays in, az_stats AND all positions out (az_engine_download_f32, 12 B per atom)
    n = n_atoms
    a = e.download()
    dpt, fpt = C.POINTER(C.c_double), C.POINTER(C.c_float)
    ne = max(1, args.e2e_steps)

    def e2e_loop(kind):
        nonlocal first
        esz = 8 if kind == "f64" else 4
        hp = lib.az_host_alloc(6 * esz * n)
        assert hp, "pinned allocation failed"
        ctype, ptype, up = (C.c_double, dpt, lib.az_engine_upload_async) if kind == "f64" else (C.c_float, fpt, lib.az_engine_upload_f32_async)
        host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(ctype)), shape=(6, n))
        host[0:3] = a["pos"].T
        host[3:6] = a["vel"].T
        ptr = [C.cast(hp + esz * n * k, ptype) for k in range(6)]
        back = None
        if kind == "f32+pos":
            bp = lib.az_host_alloc(3 * 4 * n)
            back = [C.cast(bp + 4 * n * k, fpt) for k in range(3)] + [None, None, None]
        for _ in range(2):  # warm the path (incl. the download's device buffer, allocated at its first call)
            assert up(e._h, *ptr) == 0
            e.step(first, 1); e.stats(); first += 1
            if back:
                assert lib.az_engine_download_f32(e._h, *back) == 0
        e.sync()
        t0 = time.perf_counter()
        assert up(e._h, *ptr) == 0            # pinned host memory -> staging buffer on the copy stream
        for k in range(ne):
            e.step(first, 1)                  # waits (on the device) for the staged inputs of this step
            if k + 1 < ne:
                assert up(e._h, *ptr) == 0    # next step's inputs: the copy overlaps the kernels of this step
            if back:
                assert lib.az_engine_download_f32(e._h, *back) == 0
            e.stats()                         # D2H read of this step's statistics (synchronises)
            first += 1
        t1 = time.perf_counter()
        lib.az_host_free(hp)
        if back:
            lib.az_host_free(bp)
        return {"value": n * ne / (t1 - t0), "h2d_bytes_per_step": 6 * esz * n,
                "d2h_bytes_per_step": C.sizeof(abi.az_stats) + (3 * 4 * n if back else 0)}

    e2e_f64, e2e_f32, e2e_pos = e2e_loop("f64"), e2e_loop("f32"), e2e_loop("f32+pos")
    e2e_value = e2e_f32["value"]

    out = {"metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clk,
           "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": e2e_f32["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": e2e_f32["d2h_bytes_per_step"], "steps": ne, "timer": "host wall clock",
                   "pipeline": "az_engine_upload_f32_async (float arrays, pinned): H2D of step k+1 overlaps the kernels of step k; every step re-bins and "
                               "rebuilds the list (uploaded positions invalidate it); az_stats read back every step",
                   "variants": {"f64_in_stats_out": e2e_f64, "f32_in_stats_out": e2e_f32, "f32_in_stats_and_positions_out": e2e_pos}},
           "gpu_launches": int(launches),
           "roofline": {"bound": "fp32", "kernel": dom_kernel, "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": achieved_tf / fp32_peak if fp32_peak else None, "traffic": ncu_traffic(dom_kernel, args),
                        "peak_source": "az_measure_fp32_peak in this run (FFMA chains; effective %.0f MHz x 148 SM x 128 lanes)" % mhz.value,
                        "algorithmic": "%.0f flop per unique in-cutoff pair x %d pairs per launch%s" % (
                            flop_per_pair, st.n_pairs, " + %.0f flop per atom of the integrator that rides on the kernel" % FLOP_PER_ATOM_INTEGRATE if rides else ""),
                        "avg_launch_ms": pair_ms, "launches": tm[dom]["launches"],
                        "share_of_step": tm[dom]["ms"] / ms_prof, "all_pair_kernels_share_of_step": pair_total_ms / ms_prof,
                        "limiter": (ncu_entry(dom_kernel, args) or {}).get("limiter")},
           "roofline_hbm": {"bound": "hbm", "kernels": hbm_kernels, "achieved": achieved_gbs,
                            "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                            "avg_ms": hbm_ms, "algorithmic_bytes": hbm_bytes,
                            "other_non_pair_ms_per_step": rest_ms},
           "kernel_ms_per_step": {k: v["ms"] / K for k, v in tm.items()},
           "list": {"format": "tile lists (16-bit rows into a shared-memory copy of each tile's neighbourhood)" if tile_list else "global 32-bit rows",
                    "rows_capacity": int(info.list_rows), "tile_atoms": int(info.tile_atoms), "tile_atoms_seen": int(info.tile_atoms_seen),
                    "tile_runs_seen": int(info.tile_runs_seen), "tile_rows_smem": int(info.tile_rows_smem)},
           "ns_per_day": 86400.0 / (ms / K * 1e-3) * (s.timestep * 1e3) * 1e-6,
           "section_events": {"where": "inside the timed region" if sections_inside else "separate pass of K steps after the timed region (the events add launch gaps on latency-bound decks)",
                              "ms_per_step": ms_prof / K},
           "energy": {"engKin": st.engKin_d, "engVdW": st.engVdW_d, "pairs_in_cutoff": st.n_pairs}}
    if graph:
        out["cuda_graph"] = graph
        out["config"]["cuda_graph"] = "az_engine_step replayed as one CUDA graph (value, ms_per_step)"
    e.close()
    if args.workload == "lj4m" and not args.no_small_deck:
        out["cs1"] = small_deck_line(lib)
        # the other single-GPU configurations of BASELINE.json, so that one driver-run line covers them all
        out["other_configs"] = {"cs2": small_deck_line(lib, "cs2"), "nacl1m": ionic_line()}
    rc = ref_cuda_baseline()
    if rc:
        out["ref_cuda_baseline"] = rc
    if not args.no_cpu_baseline:
        r = run_reference_sample(steps=10, warmup=2, budget_s=15.0)
        out["cpu_baseline"] = {"value": r["value"], "unit": "atom-steps/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
