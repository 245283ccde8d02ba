"""Shared test helpers: the oracle (CPU checker), the emulator build of the kernels (logic tests
without a GPU) and small synthetic systems.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from aztotmd_b200 import abi
from aztotmd_b200.engine import DescHolder, Engine, Species, System

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
EMU_DIR = os.path.join(ROOT, "tests", "emu")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)

_oracle = None
_emu = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "_build/libaz_oracle.so"], check=True)
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_build", "libaz_oracle.so"))
        lib.azo_create.restype = C.c_void_p
        lib.azo_create.argtypes = [C.POINTER(abi.az_system_desc)]
        lib.azo_destroy.argtypes = [C.c_void_p]
        lib.azo_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.azo_forces_only.argtypes = [C.c_void_p]
        lib.azo_stats.argtypes = [C.c_void_p, C.POINTER(abi.az_stats)]
        lib.azo_download.argtypes = [C.c_void_p, C.POINTER(abi.az_atoms_out), C.c_int]
        lib.azo_dump.argtypes = [C.c_void_p, C.POINTER(abi.az_dump)]
        lib.azo_force_scale.argtypes = [C.c_void_p, _fp]
        lib.azo_pairs_in_range.restype = C.c_longlong
        lib.azo_pairs_in_range.argtypes = [C.c_void_p]
        lib.azo_stencil.argtypes = [C.c_void_p, _ip]
        lib.azo_forces_allpairs.argtypes = [C.c_void_p, _fp, _fp, _fp, _dp, _dp]
        lib.azo_split_cells.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, _ip, _fp, _fp]
        lib.azo_cell_index.restype = C.c_int
        lib.azo_cell_index.argtypes = [C.c_float, C.c_float, C.c_float, _fp, _ip]
        lib.azo_pair_eval.restype = C.c_float
        lib.azo_pair_eval.argtypes = [C.c_int, _fp, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp]
        lib.azo_rdf_counts.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_ulonglong), C.c_int]
        lib.azo_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        lib.azo_default_uvects.argtypes = [_dp, _dp, _dp]
        lib.azo_angled_vector.argtypes = [_fp, C.c_float, C.c_float, _fp]
        _oracle = lib
    return _oracle


def emu_lib():
    """The kernels compiled for the host-fiber emulator (tests/emu); same C ABI as the product."""
    global _emu
    if _emu is None:
        subprocess.run(["bash", os.path.join(EMU_DIR, "build_emu.sh")], check=True, capture_output=True)
        _emu = abi.bind(C.CDLL(os.path.join(EMU_DIR, "_build", "libaztot_emu.so")), abi.ENGINE_SYMBOLS)
    return _emu


class Oracle:
    """CPU restatement of the reference's step (oracle/az_oracle.c)."""

    def rdf_counts(self, r_max, dr, nbins, npair):
        buf = np.zeros(nbins * npair, dtype=np.uint64)
        self.lib.azo_rdf_counts(self._h, r_max, dr, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), nbins)
        return buf.reshape(nbins, npair)

    def __init__(self, system):
        self.lib = oracle_lib()
        if isinstance(system, System):
            self._holder = DescHolder(system)
            desc = self._holder.desc
        else:  # a ctypes az_system_desc, e.g. from az_deck_desc
            self._holder = None
            desc = system
        self.n = desc.n_atoms
        self._h = self.lib.azo_create(C.byref(desc))
        assert self._h, "azo_create failed"
        d0 = self.dump_raw(ncell=1 << 24, probe=True)
        self.n_cells = d0

    def close(self):
        if self._h:
            self.lib.azo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, first, n=1):
        assert self.lib.azo_step(self._h, first, n) == 0

    def forces_only(self):
        assert self.lib.azo_forces_only(self._h) == 0

    def stats(self):
        st = abi.az_stats()
        assert self.lib.azo_stats(self._h, C.byref(st)) == 0
        return st

    def dump_raw(self, ncell, probe=False):
        n = self.n
        d = abi.az_dump()
        d.n_atoms, d.n_cells = n, ncell
        if probe:
            assert self.lib.azo_dump(self._h, C.byref(d)) == 0
            return d.n_cells
        return d

    def dump(self):
        n, nc = self.n, self.n_cells
        d = abi.az_dump()
        d.n_atoms, d.n_cells = n, nc
        ints = {k: np.zeros(m, dtype=np.int32) for k, m in (("cell_indexes", n), ("first_atom_in_cell", nc), ("n_at_in_cell", nc), ("ids", n))}
        for k, a in ints.items():
            setattr(d, k, a.ctypes.data_as(_ip))
        fl = {k: np.zeros(n, dtype=np.float32) for k in ("fx", "fy", "fz", "x", "y", "z", "vx", "vy", "vz")}
        for k, a in fl.items():
            setattr(d, k, a.ctypes.data_as(_fp))
        assert self.lib.azo_dump(self._h, C.byref(d)) == 0
        return dict(cnum=tuple(d.cnum), csize=tuple(d.csize), crev=tuple(d.crev), cell=ints["cell_indexes"],
                    first=ints["first_atom_in_cell"], count=ints["n_at_in_cell"], ids=ints["ids"],
                    frc=np.stack([fl["fx"], fl["fy"], fl["fz"]], 1), pos=np.stack([fl["x"], fl["y"], fl["z"]], 1),
                    vel=np.stack([fl["vx"], fl["vy"], fl["vz"]], 1))

    def download(self, order=abi.ORDER_ORIGINAL):
        n = self.n
        out = abi.az_atoms_out()
        out.n_atoms = n
        arr = {k: np.zeros(n, dtype=np.float64) for k in ("xs", "ys", "zs", "vxs", "vys", "vzs")}
        for k, a in arr.items():
            setattr(out, k, a.ctypes.data_as(_dp))
        types = np.zeros(n, dtype=np.int32); ids = np.zeros(n, dtype=np.int32)
        engs = np.zeros(n, dtype=np.float32); radii = np.zeros(n, dtype=np.float32)
        out.types = types.ctypes.data_as(_ip); out.ids = ids.ctypes.data_as(_ip)
        out.engs = engs.ctypes.data_as(_fp); out.radii = radii.ctypes.data_as(_fp)
        assert self.lib.azo_download(self._h, C.byref(out), order) == 0
        return dict(pos=np.stack([arr["xs"], arr["ys"], arr["zs"]], 1), vel=np.stack([arr["vxs"], arr["vys"], arr["vzs"]], 1),
                    types=types, ids=ids, engs=engs, radii=radii)

    def force_scale(self):
        a = np.zeros(self.n, dtype=np.float32)
        self.lib.azo_force_scale(self._h, a.ctypes.data_as(_fp))
        return a

    def pairs_in_range(self):
        return int(self.lib.azo_pairs_in_range(self._h))

    def stencil(self):
        a = np.zeros(3, dtype=np.int32)
        self.lib.azo_stencil(self._h, a.ctypes.data_as(_ip))
        return tuple(int(x) for x in a)

    def forces_allpairs(self):
        n = self.n
        f = [np.zeros(n, dtype=np.float32) for _ in range(3)]
        ev, ec = C.c_double(), C.c_double()
        self.lib.azo_forces_allpairs(self._h, *[a.ctypes.data_as(_fp) for a in f], C.byref(ev), C.byref(ec))
        return np.stack(f, 1), ev.value, ec.value


# ----------------------------------------------------------------------------- systems
AR_EPS, AR_SIGMA, AR_MASS = 0.01006, 3.3952, 39.9


def fcc(ncell, a, jitter, seed):
    from aztotmd_b200.decks import fcc_positions
    return fcc_positions(ncell, a, jitter=jitter, seed=seed)


def lj_system(ncell=(4, 4, 4), rho_star=0.8442, rc_sigma=2.5, temperature=86.6, seed=1, cell_size=None,
              jitter=0.05, vel=True, tstat=abi.TSTAT_NONE, n_equil=0, eq_freq=0):
    a = (4.0 / rho_star) ** (1.0 / 3.0) * AR_SIGMA
    pos, box = fcc(ncell, a, jitter, seed)
    rc = rc_sigma * AR_SIGMA
    s = System(pos=pos, types=np.zeros(len(pos), dtype=np.int32), box=box, species=[Species("Ar", AR_MASS)],
               timestep=0.001, cell_size=cell_size if cell_size else rc, temperature=temperature, seed=seed,
               tstat_kind=tstat, n_equil=n_equil, eq_freq=eq_freq)
    s.set_vdw(0, 0, "lnjs", rc, [AR_EPS, AR_SIGMA])
    if vel:
        s.vel = s.gauss_velocities(seed + 100)
    return s


def nacl_system(ncell=(3, 3, 3), a=5.9, rc_vdw=9.0, r_real=9.0, alpha=0.2, elec=abi.ELEC_FENNEL, temperature=1200.0,
                seed=2, cell_size=None, vel=True, kind="buck", jitter=0.05, ewald_k=(0, 0, 0)):
    from aztotmd_b200.decks import rocksalt_positions
    pos, types, box = rocksalt_positions(ncell, a, jitter=jitter, seed=seed)
    s = System(pos=pos, types=types.astype(np.int32), box=box,
               species=[Species("Na", 22.99, 1.0), Species("Cl", 35.45, -1.0)], timestep=0.001,
               cell_size=cell_size if cell_size else max(r_real, rc_vdw), elec_kind=elec, r_real=r_real, alpha=alpha,
               temperature=temperature, seed=seed, ewald_k=ewald_k)
    if kind == "buck":
        s.set_vdw(0, 0, "buck", rc_vdw, [424.097, 0.317, 1.05])
        s.set_vdw(0, 1, "buck", rc_vdw, [1256.31, 0.317, 7.00])
        s.set_vdw(1, 1, "buck", rc_vdw, [3488.998, 0.317, 72.40])
    elif kind == "mixed":  # exercises the per-pair switch: three different kinds in one table
        s.set_vdw(0, 0, "lnjs", rc_vdw, [0.005, 2.5])
        s.set_vdw(0, 1, "buck", rc_vdw, [1256.31, 0.317, 7.00])
        s.set_vdw(1, 1, "bmhs", rc_vdw, [0.25, 3.0, 3.2, 70.0, 140.0])
    elif kind == "mixed2":  # the remaining analytic forms: elin, einv, bmhs
        s.set_vdw(0, 0, "elin", rc_vdw, [900.0, 0.33, 0.02])
        s.set_vdw(0, 1, "einv", rc_vdw, [900.0, 0.33, 1.5])
        s.set_vdw(1, 1, "bmhs", rc_vdw, [0.25, 3.0, 3.2, 70.0, 140.0])
    if vel:
        s.vel = s.gauss_velocities(seed + 100)
    return s


def surk_system(n_side=6, spacing=3.0, box=35.0, temperature=500.0, seed=3, cell_size=2.7, tstat=abi.TSTAT_RADI,
                n_equil=0, eq_freq=0, vel=True):
    """A droplet of 'abstract atoms' like case study 2: surk 6.0 75.0 8.0 1.0 1.0 + radii A=2.73 B=4.731 Umax=0.2."""
    rng = np.random.default_rng(seed)
    g = np.arange(n_side) * spacing
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pos = pos + (box - g[-1]) / 2 + rng.uniform(-0.2, 0.2, size=pos.shape)
    s = System(pos=pos, types=np.zeros(len(pos), dtype=np.int32), box=(box, box, box),
               species=[Species("Ar", 39.9, 0.0, 0, 2.73, 4.731, 0.2)], timestep=0.001, cell_size=cell_size,
               temperature=temperature, seed=seed, tstat_kind=tstat, n_equil=n_equil, eq_freq=eq_freq)
    s.set_vdw(0, 0, "surk", 6.0, [75.0, 8.0, 1.0, 1.0])
    if vel:
        s.vel = s.gauss_velocities(seed + 100)
    return s


def sort_by_id(d):
    """Canonical view of a sorted-order dump: reorder per-atom arrays by persistent id."""
    order = np.argsort(d["ids"], kind="stable")
    return {k: (v[order] if isinstance(v, np.ndarray) and len(v) == len(order) and k not in ("first", "count") else v)
            for k, v in d.items()}
