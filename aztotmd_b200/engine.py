"""Python host side of the B200 engine: system description + thin wrapper over the C ABI.

``System`` mirrors what the reference's host layer (``init_md`` src/sys_init.cpp:1036-1119,
``read_vdw`` src/vdw.cpp:234-308, ``prepare_elec`` src/elec.cpp:371-406) hands to
``init_cudaMD``; ``Engine`` calls ``az_engine_*`` of include/aztot_b200.h.  All compute happens in
the CUDA library: if it is missing, or no GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import abi

# unit factors of the reference, src/const.h:16-49
E_SI = 1.60217733e-19
M_SCALE = 1.6605402e-27 / (E_SI * 1.0e-12 * 1.0e-12 / 1.0e-10 / 1.0e-10)  # amu -> internal mass
KB = 1.3806488e-23 / 1.60217733e-19  # eV/K
PI = 3.14159265359  # src/const.h:11 (not math.pi: the reference's truncated literal)
SQRTPI = math.sqrt(PI)


class AzError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s: %s" % (abi.STATUS_NAMES.get(status, str(status)), message))
        self.status = status


@dataclass
class Species:
    name: str
    mass_amu: float
    charge: float = 0.0
    frozen: int = 0
    radA: float = 0.0
    radB: float = 0.0
    mxEng: float = 0.0


def prepare_vdw(kind: str, rcut: float, params: Sequence[float]) -> dict:
    """One 'vdw' line -> prepared entry, following read_vdw (src/vdw.cpp:261-299); r/E scales are 1."""
    k = abi.VDW_BY_NAME[kind]
    p = list(params) + [0.0] * (5 - len(params))
    if k == abi.VDW_LJ:  # vdw_scale0 = 4: p0 = 4e; p1 = s^2; p2 = 6*p0 = 24e
        p[0] = 4.0 * p[0]
        p[1] = p[1] * p[1]
        p[2] = 6.0 * p[0]
    return dict(kind=k, use_radii=1 if k == abi.VDW_SURK else 0, p=p, r2cut=rcut * rcut, rcut=rcut)


@dataclass
class System:
    """Everything az_system_desc carries, in Python terms (positions in Angstrom, time in ps)."""
    pos: np.ndarray                 # [N,3] float64
    types: np.ndarray               # [N] int32
    box: Sequence[float]            # la, lb, lc
    species: Sequence[Species]
    vdw: dict = field(default_factory=dict)   # (i,j) -> prepare_vdw(...) entry; symmetric fill like read_vdw
    vel: Optional[np.ndarray] = None
    timestep: float = 0.001
    cell_size: float = 0.0
    elec_kind: int = abi.ELEC_NONE
    r_real: float = 0.0
    alpha: float = 0.0
    eps: float = 1.0
    tstat_kind: int = abi.TSTAT_NONE
    temperature: float = 0.0
    n_equil: int = 0
    eq_freq: int = 0
    efield: Sequence[float] = (0.0, 0.0, 0.0)
    photons: Optional[np.ndarray] = None
    init_radii: Optional[np.ndarray] = None
    init_engs: Optional[np.ndarray] = None
    seed: int = 12345
    # constant bonds / valence angles.  bond_types[k] = (kind name, params), k = 1.. (type 0 is the empty entry);
    # bonds = rows (atom 1, atom 2, type); angle_types[k] = (k, cos theta0); angles = rows (centre, ligand 1, ligand 2, type)
    bond_types: Sequence = ()
    bonds: Optional[np.ndarray] = None
    angle_types: Sequence = ()
    angles: Optional[np.ndarray] = None
    # domain decomposition, distributed input: this rank's share of the system with its persistent ids
    global_ids: Optional[np.ndarray] = None
    n_atoms_global: int = 0
    tstat_tau: float = 0.0   # Nose-Hoover relaxation time, ps ('temperature T nose tau')
    reset_vels: int = 0      # every n-th step: zero the velocities instead of the second half kick
    ewald_k: Sequence[int] = (0, 0, 0)  # 'elec pme rc alpha kx ky kz': reciprocal-space Ewald when any is > 0

    def set_vdw(self, i: int, j: int, kind: str, rcut: float, params: Sequence[float]):
        e = prepare_vdw(kind, rcut, params)
        self.vdw[(i, j)] = e
        if e["kind"] != abi.VDW_SURK:
            self.vdw[(j, i)] = e  # src/vdw.cpp:306-307
        elif i != j:
            # the reference leaves [j][i] empty (asymmetric, src/vdw.cpp:292-298,306); a full
            # neighbour list needs the mirrored entry with k and l swapped so both sides agree
            m = dict(e)
            m["p"] = [e["p"][0], e["p"][1], e["p"][3], e["p"][2], e["p"][4]]
            self.vdw[(j, i)] = m

    # ---- derived quantities exactly as init_md computes them
    @property
    def max_r_vdw(self) -> float:
        return max([v["rcut"] for v in self.vdw.values()], default=0.0)

    @property
    def charged(self) -> bool:
        return any(s.charge != 0.0 for s in self.species)

    @property
    def effective_elec_kind(self) -> int:
        # read_elec downgrades to 'none' when no species is charged (src/elec.cpp:52-57)
        return self.elec_kind if self.charged else abi.ELEC_NONE

    @property
    def r_max(self) -> float:  # src/sys_init.cpp:1059-1072
        return self.r_real if self.effective_elec_kind else self.max_r_vdw

    @property
    def n_bonds(self) -> int:
        return 0 if self.bonds is None else len(self.bonds)

    @property
    def deg_free(self) -> int:  # src/sys_init.cpp:1099-1102
        n = self.n_atoms_global if self.n_atoms_global > 0 else len(self.pos)
        return 3 * n - self.n_bonds - (1 if self.tstat_kind else 0)

    @property
    def t_kin(self) -> float:  # src/sys_init.cpp:1106
        return 0.5 * self.temperature * KB * self.deg_free

    def gauss_velocities(self, seed: int) -> np.ndarray:
        """Maxwell velocities at `temperature` with zero total momentum (what 'init_vel gaus' asks for;
        the reference draws them from the time-seeded rand(), src/md_utils.cpp:7-26)."""
        rng = np.random.default_rng(seed)
        m = np.array([s.mass_amu * M_SCALE for s in self.species])[self.types]
        v = rng.standard_normal((len(self.pos), 3)) * np.sqrt(KB * self.temperature / m)[:, None]
        v -= (v * m[:, None]).sum(0) / m.sum()
        ke = 0.5 * (m[:, None] * v * v).sum()
        if ke > 0:
            v *= math.sqrt(0.5 * KB * self.temperature * 3 * len(self.pos) / ke)
        return v


def photon_energies(n: int, temperature: float, seed: int) -> np.ndarray:
    """Photon energies with density ~ e^4 exp(-e/kT) on 0..1 eV (what photon_engs samples by
    bisection of a series CDF from rand01(), src/temperature.cpp:28-89).  Sampled here as a
    Gamma(5, kT) variate truncated to 1 eV; seeded, unlike the reference."""
    rng = np.random.default_rng(seed)
    kt = KB * temperature
    out = rng.gamma(5.0, kt, size=n)
    bad = out > 1.0
    while bad.any():
        out[bad] = rng.gamma(5.0, kt, size=int(bad.sum()))
        bad = out > 1.0
    return out


class DescHolder:
    """Builds az_system_desc and keeps every referenced numpy buffer alive."""

    def __init__(self, s: System):
        n = len(s.pos)
        self.n = n
        self.keep = []

        def dbl(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self.keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double))

        d = abi.az_system_desc()
        d.abi_version = abi.AZ_ABI_VERSION
        d.n_atoms = n
        d.n_species = len(s.species)
        pos = np.asarray(s.pos, dtype=np.float64)
        d.xs, d.ys, d.zs = dbl(pos[:, 0]), dbl(pos[:, 1]), dbl(pos[:, 2])
        if s.vel is not None:
            vel = np.asarray(s.vel, dtype=np.float64)
            d.vxs, d.vys, d.vzs = dbl(vel[:, 0]), dbl(vel[:, 1]), dbl(vel[:, 2])
        types = np.ascontiguousarray(s.types, dtype=np.int32)
        self.keep.append(types)
        d.types = types.ctypes.data_as(C.POINTER(C.c_int))
        d.la, d.lb, d.lc = [float(x) for x in s.box]
        ns = len(s.species)
        sp = (abi.az_species * ns)()
        for i, x in enumerate(s.species):
            sp[i].name = x.name.encode()[:7]
            sp[i].mass = x.mass_amu * M_SCALE  # src/sys_init.cpp read_spec: mass *= m_scale
            sp[i].charge = x.charge
            sp[i].frozen = x.frozen
            sp[i].radA, sp[i].radB, sp[i].mxEng = x.radA, x.radB, x.mxEng
        self.keep.append(sp)
        d.species = sp
        tab = (abi.az_vdw_entry * (ns * ns))()
        for (i, j), e in s.vdw.items():
            t = tab[i * ns + j]
            t.kind, t.use_radii = e["kind"], e["use_radii"]
            t.p0, t.p1, t.p2, t.p3, t.p4 = e["p"]
            t.r2cut = e["r2cut"]
        self.keep.append(tab)
        d.vdw = tab
        d.max_r_vdw = s.max_r_vdw
        ek = s.effective_elec_kind
        d.elec_kind = ek
        d.r_real, d.alpha, d.eps = s.r_real, s.alpha, s.eps
        if ek in (abi.ELEC_EWALD, abi.ELEC_FENNEL):
            d.daipi2 = 2.0 * s.alpha / SQRTPI
        if ek == abi.ELEC_FENNEL:  # src/elec.cpp:399-405
            arc = s.alpha * s.r_real
            d.el_c1 = math.erfc(arc) / s.r_real
            d.el_c2 = math.erfc(arc) / (s.r_real * s.r_real) + d.daipi2 * math.exp(-arc * arc) / s.r_real
        d.tstat_kind = s.tstat_kind
        d.temperature = s.temperature
        d.t_kin = s.t_kin
        if s.tstat_kind == abi.TSTAT_RADI:
            ph = s.photons if s.photons is not None else photon_energies(n, s.temperature, s.seed)
            d.photons = dbl(ph)
        if s.init_radii is not None:
            a = np.ascontiguousarray(s.init_radii, dtype=np.float32)
            self.keep.append(a)
            d.init_radii = a.ctypes.data_as(C.POINTER(C.c_float))
        if s.init_engs is not None:
            a = np.ascontiguousarray(s.init_engs, dtype=np.float32)
            self.keep.append(a)
            d.init_engs = a.ctypes.data_as(C.POINTER(C.c_float))
        d.seed = s.seed
        d.timestep = s.timestep
        d.n_equil, d.eq_freq = s.n_equil, s.eq_freq
        d.r2_max = s.r_max * s.r_max
        d.cell_size = s.cell_size
        for a in range(3):
            d.efield[a] = float(s.efield[a])
        if s.global_ids is not None:
            gi = np.ascontiguousarray(s.global_ids, dtype=np.int32)
            self.keep.append(gi)
            d.global_ids = gi.ctypes.data_as(C.POINTER(C.c_int))
        d.n_atoms_global = int(s.n_atoms_global)
        d.tstat_tau = float(s.tstat_tau)
        d.reset_vels = int(s.reset_vels)
        for a in range(3):
            d.ewald_k[a] = int(s.ewald_k[a]) if ek == abi.ELEC_EWALD else 0

        def ints(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            self.keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int))

        if s.bonds is not None and len(s.bonds):
            bt = (abi.az_bond_type * (len(s.bond_types) + 1))()
            for k, (name, par) in enumerate(s.bond_types):
                t = bt[k + 1]
                t.kind = abi.BOND_BY_NAME[name]
                p = list(par) + [0.0] * (5 - len(par))
                t.p0, t.p1, t.p2, t.p3, t.p4 = p
            self.keep.append(bt)
            b = np.asarray(s.bonds, dtype=np.int32)
            d.n_bond_types, d.n_bonds, d.bond_types = len(bt), len(b), bt
            d.bond_at1, d.bond_at2, d.bond_type = ints(b[:, 0]), ints(b[:, 1]), ints(b[:, 2])
        if s.angles is not None and len(s.angles):
            at = (abi.az_angle_type * (len(s.angle_types) + 1))()
            for k, (kk, c0) in enumerate(s.angle_types):
                at[k + 1].kind, at[k + 1].p0, at[k + 1].p1 = abi.ANGLE_HCOS, kk, c0
            self.keep.append(at)
            a = np.asarray(s.angles, dtype=np.int32)
            d.n_angle_types, d.n_angles, d.angle_types = len(at), len(a), at
            d.angle_centre, d.angle_lig1, d.angle_lig2, d.angle_type = ints(a[:, 0]), ints(a[:, 1]), ints(a[:, 2]), ints(a[:, 3])
        self.desc = d


def unique_comm_id(lib=None) -> bytes:
    """128 bytes that identify one multi-GPU job (az_comm_unique_id); rank 0 creates them and every
    rank passes the same bytes to Engine(..., comm_id=...)."""
    lib = lib if lib is not None else abi.load_library()
    buf = C.create_string_buffer(128)
    rc = lib.az_comm_unique_id(buf)
    if rc != abi.AZ_OK:
        msg = lib.az_last_error(None)
        raise AzError(rc, msg.decode() if msg else "")
    return buf.raw


class Engine:
    """One engine = one GPU.  ``lib`` defaults to the nvcc-built library (no fallback)."""

    def __init__(self, system, *, device: int = 0, rebuild_every: int = 1, skin: float = 0.0,
                 cells_per_block: int = 0, pair_variant: int = 0, max_neigh: int = 0, use_graph: bool = False, lib=None,
                 list_format: int = 0, tile_atoms: int = 0,
                 opts: Optional[abi.az_cuda_opts] = None,
                 rank: int = 0, n_ranks: int = 1, pgrid: Sequence[int] = (0, 0, 0), comm_id: Optional[bytes] = None):
        self.lib = lib if lib is not None else abi.load_library()
        if isinstance(system, System):
            self._holder = DescHolder(system)
            desc = self._holder.desc
        else:  # a ctypes az_system_desc (e.g. from az_deck_desc)
            self._holder = None
            desc = system
        self.n = desc.n_atoms
        self.nspec = desc.n_species
        if opts is None:
            opts = abi.az_cuda_opts()
            opts.device = device
            opts.rebuild_every = rebuild_every
            opts.skin = skin
            opts.cells_per_block = cells_per_block
            opts.pair_variant = pair_variant
            opts.max_neigh = max_neigh
            opts.list_format, opts.tile_atoms = list_format, tile_atoms
            opts.use_graph = 1 if use_graph else 0
            opts.rank, opts.n_ranks = rank, n_ranks
            for a in range(3):
                opts.pgrid[a] = int(pgrid[a])
            if n_ranks > 1:
                if comm_id is None or len(comm_id) != 128:
                    raise ValueError("n_ranks > 1 needs comm_id: the 128 bytes of unique_comm_id(), the same on every rank")
                self._comm_id = C.create_string_buffer(bytes(comm_id), 128)
                opts.nccl_id = C.cast(self._comm_id, C.c_void_p)
        self._h = C.c_void_p()
        rc = self.lib.az_engine_create(C.byref(desc), C.byref(opts), C.byref(self._h))
        if rc != abi.AZ_OK:
            msg = self.lib.az_last_error(None)
            raise AzError(rc, msg.decode() if msg else "")
        self.info = abi.az_engine_info_t()
        self._check(self.lib.az_engine_info(self._h, C.byref(self.info)))
        self.n_cells = self.info.n_cells
        self.n = max(self.n, self.info.capacity)  # a brick may come to own more atoms than it was given

    def _check(self, rc: int):
        if rc != abi.AZ_OK:
            msg = self.lib.az_last_error(self._h)
            raise AzError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.az_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def refresh_info(self) -> abi.az_engine_info_t:
        self._check(self.lib.az_engine_info(self._h, C.byref(self.info)))
        return self.info

    def step(self, first_step: int, nsteps: int = 1):
        self._check(self.lib.az_engine_step(self._h, first_step, nsteps))

    def step_timed(self, first_step: int, nsteps: int) -> float:
        """Advance nsteps; returns the device time in ms (CUDA events on the engine's stream)."""
        ms = C.c_float()
        self._check(self.lib.az_engine_step_timed(self._h, first_step, nsteps, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._check(self.lib.az_engine_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.az_engine_stream(self._h) or 0)

    def set_option(self, key: str, value: float):
        self._check(self.lib.az_engine_set_option(self._h, key.encode(), float(value)))

    def stats(self) -> abi.az_stats:
        st = abi.az_stats()
        self._check(self.lib.az_engine_stats(self._h, C.byref(st)))
        return st

    def download(self, order: int = abi.ORDER_ORIGINAL) -> dict:
        n = self.n
        out = abi.az_atoms_out()
        out.n_atoms = n
        arr = {k: np.zeros(n, dtype=np.float64) for k in ("xs", "ys", "zs", "vxs", "vys", "vzs")}
        for k, a in arr.items():
            setattr(out, k, a.ctypes.data_as(C.POINTER(C.c_double)))
        types = np.zeros(n, dtype=np.int32)
        ids = np.zeros(n, dtype=np.int32)
        engs = np.zeros(n, dtype=np.float32)
        radii = np.zeros(n, dtype=np.float32)
        out.types = types.ctypes.data_as(C.POINTER(C.c_int))
        out.ids = ids.ctypes.data_as(C.POINTER(C.c_int))
        out.engs = engs.ctypes.data_as(C.POINTER(C.c_float))
        out.radii = radii.ctypes.data_as(C.POINTER(C.c_float))
        self._check(self.lib.az_engine_download(self._h, C.byref(out), order))
        k = out.n_written
        return dict(pos=np.stack([arr["xs"], arr["ys"], arr["zs"]], 1)[:k], vel=np.stack([arr["vxs"], arr["vys"], arr["vzs"]], 1)[:k],
                    types=types[:k], ids=ids[:k], engs=engs[:k], radii=radii[:k])

    def dump(self) -> dict:
        n, nc = self.n, self.n_cells
        d = abi.az_dump()
        d.n_atoms, d.n_cells = n, nc
        ints = {k: np.zeros(m, dtype=np.int32) for k, m in (("cell_indexes", n), ("first_atom_in_cell", nc), ("n_at_in_cell", nc), ("ids", n))}
        for k, a in ints.items():
            setattr(d, k, a.ctypes.data_as(C.POINTER(C.c_int)))
        fl = {k: np.zeros(n, dtype=np.float32) for k in ("fx", "fy", "fz", "x", "y", "z", "vx", "vy", "vz")}
        for k, a in fl.items():
            setattr(d, k, a.ctypes.data_as(C.POINTER(C.c_float)))
        self._check(self.lib.az_engine_debug_dump(self._h, C.byref(d)))
        n2, nc2 = d.n_atoms, d.n_cells
        return dict(cnum=tuple(d.cnum), csize=tuple(d.csize), crev=tuple(d.crev),
                    cell=ints["cell_indexes"][:n2], first=ints["first_atom_in_cell"][:nc2], count=ints["n_at_in_cell"][:nc2],
                    ids=ints["ids"][:n2],
                    frc=np.stack([fl["fx"], fl["fy"], fl["fz"]], 1)[:n2], pos=np.stack([fl["x"], fl["y"], fl["z"]], 1)[:n2],
                    vel=np.stack([fl["vx"], fl["vy"], fl["vz"]], 1)[:n2])

    def upload(self, pos: np.ndarray, vel: Optional[np.ndarray] = None, asynchronous: bool = False):
        cols = [np.ascontiguousarray(pos[:, a], dtype=np.float64) for a in range(3)]
        vcols = [np.ascontiguousarray(vel[:, a], dtype=np.float64) for a in range(3)] if vel is not None else [None] * 3
        ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
        fn = self.lib.az_engine_upload_async if asynchronous else self.lib.az_engine_upload
        if asynchronous:
            self._pending_upload = (cols, vcols)  # the arrays must outlive the copy
        self._check(fn(self._h, *[ptr(a) for a in cols], *[ptr(a) for a in vcols]))

    def upload_f32(self, pos: np.ndarray, vel: Optional[np.ndarray] = None):
        """az_engine_upload_f32_async: float arrays, half the host->device bytes; asynchronous (the arrays are kept alive)."""
        cols = [np.ascontiguousarray(pos[:, a], dtype=np.float32) for a in range(3)]
        vcols = [np.ascontiguousarray(vel[:, a], dtype=np.float32) for a in range(3)] if vel is not None else [None] * 3
        ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None
        self._pending_upload = (cols, vcols)
        self._check(self.lib.az_engine_upload_f32_async(self._h, *[ptr(a) for a in cols], *[ptr(a) for a in vcols]))

    def download_f32(self, with_vel: bool = True) -> dict:
        """az_engine_download_f32: positions (wrapped) and velocities in atoms.xyz order as float32."""
        n = self.n
        a = [np.zeros(n, dtype=np.float32) for _ in range(6 if with_vel else 3)]
        ptr = [x.ctypes.data_as(C.POINTER(C.c_float)) for x in a] + [None] * (0 if with_vel else 3)
        self._check(self.lib.az_engine_download_f32(self._h, *ptr))
        out = dict(pos=np.stack(a[0:3], 1))
        if with_vel:
            out["vel"] = np.stack(a[3:6], 1)
        return out

    def rdf_begin(self, r_max: float, dr: float):
        self._check(self.lib.az_engine_rdf_begin(self._h, float(r_max), float(dr)))

    def rdf_sample(self):
        self._check(self.lib.az_engine_rdf_sample(self._h))

    def rdf_counts(self):
        """(counts[n_bins, n_pairs] of unordered pairs per bin and species pair, n_samples)."""
        nb, npair, ns = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.az_engine_rdf_read(self._h, None, 0, C.byref(nb), C.byref(npair), C.byref(ns)))
        buf = np.zeros(nb.value * npair.value, dtype=np.uint64)
        self._check(self.lib.az_engine_rdf_read(self._h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.size, None, None, None))
        return buf.reshape(nb.value, npair.value), ns.value

    def profile(self, mode: int = 1):
        """0 off, 1 synchronous per-section timing, 2 recorded events (no host sync in the loop)."""
        self._check(self.lib.az_engine_profile(self._h, int(mode)))

    def timers(self) -> dict:
        ms = (C.c_double * abi.AZ_N_TIMERS)()
        ln = (C.c_longlong * abi.AZ_N_TIMERS)()
        self._check(self.lib.az_engine_timers(self._h, ms, ln))
        return {nm: dict(ms=ms[k], launches=ln[k]) for k, nm in enumerate(abi.TIMER_NAMES)}

    @property
    def launch_count(self) -> int:
        return int(self.lib.az_engine_launch_count(self._h))
