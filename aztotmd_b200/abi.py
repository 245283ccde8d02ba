"""ctypes mirror of include/aztot_b200.h (the C ABI of the B200 engine).

The structures below must stay field-for-field identical to the header; tests/test_abi.py
checks sizes and every exported symbol.  ``load_library()`` loads the nvcc-built
``aztotmd_b200/lib/libaztot_b200.so`` and raises if it is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

AZ_ABI_VERSION = 2
AZ_MAX_SPECIES = 15
AZ_N_UVECT = 3072
AZ_N_MOM = 20
AZ_MX_FRC = 1.0e4
AZ_N_TIMERS = 9

AZ_OK = 0
AZ_ERR_INVALID, AZ_ERR_NO_DEVICE, AZ_ERR_CUDA, AZ_ERR_UNSUPPORTED = -1, -2, -3, -4
AZ_ERR_OVERFLOW, AZ_ERR_COMM, AZ_ERR_IO = -5, -6, -7
STATUS_NAMES = {0: "AZ_OK", -1: "AZ_ERR_INVALID", -2: "AZ_ERR_NO_DEVICE", -3: "AZ_ERR_CUDA",
                -4: "AZ_ERR_UNSUPPORTED", -5: "AZ_ERR_OVERFLOW", -6: "AZ_ERR_COMM", -7: "AZ_ERR_IO"}

VDW_NONE, VDW_LJ, VDW_BUCK, VDW_P746, VDW_BMH, VDW_ELIN, VDW_EINV, VDW_SURK = range(8)
VDW_BY_NAME = {"lnjs": VDW_LJ, "buck": VDW_BUCK, "p746": VDW_P746, "bmhs": VDW_BMH, "elin": VDW_ELIN,
               "einv": VDW_EINV, "surk": VDW_SURK}
ELEC_NONE, ELEC_DIRECT, ELEC_EWALD, ELEC_FENNEL = range(4)
TSTAT_NONE, TSTAT_NOSE, TSTAT_RADI = range(3)
ORDER_SORTED, ORDER_ORIGINAL = 0, 1
TIMER_NAMES = ["integrate_a", "scan", "sort", "pair", "integrate_b", "halo", "pair_list", "ewald", "migrate"]

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


class az_species(C.Structure):
    _fields_ = [("name", C.c_char * 8), ("mass", C.c_double), ("charge", C.c_double), ("frozen", C.c_int),
                ("reserved", C.c_int), ("radA", C.c_double), ("radB", C.c_double), ("mxEng", C.c_double)]


class az_vdw_entry(C.Structure):
    _fields_ = [("kind", C.c_int), ("use_radii", C.c_int), ("p0", C.c_double), ("p1", C.c_double),
                ("p2", C.c_double), ("p3", C.c_double), ("p4", C.c_double), ("r2cut", C.c_double)]


BOND_NONE, BOND_HARM, BOND_MORSE, BOND_PEDONE, BOND_BUCK, BOND_E612 = range(6)
BOND_BY_NAME = {"harm": BOND_HARM, "mors": BOND_MORSE, "pdn": BOND_PEDONE, "buck": BOND_BUCK, "e612": BOND_E612}
ANGLE_HCOS = 1


class az_bond_type(C.Structure):
    _fields_ = [("kind", C.c_int), ("spec1", C.c_int), ("spec2", C.c_int), ("reserved", C.c_int),
                ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double), ("p3", C.c_double), ("p4", C.c_double)]


class az_angle_type(C.Structure):
    _fields_ = [("kind", C.c_int), ("central", C.c_int), ("p0", C.c_double), ("p1", C.c_double)]


class az_system_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("n_atoms", C.c_int), ("n_species", C.c_int), ("reserved0", C.c_int),
        ("xs", _dp), ("ys", _dp), ("zs", _dp), ("vxs", _dp), ("vys", _dp), ("vzs", _dp),
        ("fxs", _dp), ("fys", _dp), ("fzs", _dp), ("types", _ip),
        ("la", C.c_double), ("lb", C.c_double), ("lc", C.c_double),
        ("species", C.POINTER(az_species)), ("vdw", C.POINTER(az_vdw_entry)), ("max_r_vdw", C.c_double),
        ("elec_kind", C.c_int), ("reserved1", C.c_int),
        ("r_real", C.c_double), ("alpha", C.c_double), ("daipi2", C.c_double), ("el_c1", C.c_double),
        ("el_c2", C.c_double), ("eps", C.c_double),
        ("tstat_kind", C.c_int), ("reserved2", C.c_int),
        ("temperature", C.c_double), ("t_kin", C.c_double),
        ("photons", _dp), ("uvect_x", _dp), ("uvect_y", _dp), ("uvect_z", _dp),
        ("init_radii", _fp), ("init_engs", _fp), ("seed", C.c_uint64),
        ("timestep", C.c_double), ("n_equil", C.c_int), ("eq_freq", C.c_int),
        ("r2_max", C.c_double), ("cell_size", C.c_double), ("efield", C.c_double * 3),
        ("global_ids", _ip), ("n_atoms_global", C.c_longlong),
        ("tstat_tau", C.c_double), ("reset_vels", C.c_int), ("reserved3", C.c_int),
        ("ewald_k", C.c_int * 3), ("reserved4", C.c_int),
        ("n_bond_types", C.c_int), ("n_bonds", C.c_int), ("bond_types", C.POINTER(az_bond_type)),
        ("bond_at1", _ip), ("bond_at2", _ip), ("bond_type", _ip),
        ("n_angle_types", C.c_int), ("n_angles", C.c_int), ("angle_types", C.POINTER(az_angle_type)),
        ("angle_centre", _ip), ("angle_lig1", _ip), ("angle_lig2", _ip), ("angle_type", _ip),
    ]


class az_cuda_opts(C.Structure):
    _fields_ = [("device", C.c_int), ("rebuild_every", C.c_int), ("skin", C.c_float), ("use_graph", C.c_int),
                ("pair_variant", C.c_int), ("cells_per_block", C.c_int), ("rank", C.c_int), ("n_ranks", C.c_int),
                ("pgrid", C.c_int * 3), ("ghost_slots", C.c_int), ("nccl_id", C.c_void_p),
                ("max_neigh", C.c_int), ("list_format", C.c_int), ("tile_atoms", C.c_int), ("reserved", C.c_int * 5)]


class az_stats(C.Structure):
    _fields_ = [("step", C.c_int), ("n_atoms_total", C.c_int),
                ("engTot", C.c_float), ("engKin", C.c_float), ("engVdW", C.c_float), ("engCoul1", C.c_float),
                ("engCoul2", C.c_float), ("engTemp", C.c_float), ("engElecField", C.c_float), ("engPot", C.c_float),
                ("posMom", C.c_float * 3), ("negMom", C.c_float * 3), ("pressure", C.c_float), ("x_first", C.c_float),
                ("specAcBoxPos", (C.c_int * 3) * AZ_MAX_SPECIES), ("specAcBoxNeg", (C.c_int * 3) * AZ_MAX_SPECIES),
                ("engKin_d", C.c_double), ("engVdW_d", C.c_double), ("engCoul1_d", C.c_double), ("engTemp_d", C.c_double),
                ("n_pairs", C.c_longlong), ("engCoul3", C.c_float), ("reserved", C.c_float), ("engCoul2_d", C.c_double),
                ("engBond", C.c_float), ("engAngle", C.c_float), ("engBond_d", C.c_double), ("engAngle_d", C.c_double)]


class az_atoms_out(C.Structure):
    _fields_ = [("n_atoms", C.c_int), ("n_written", C.c_int), ("xs", _dp), ("ys", _dp), ("zs", _dp),
                ("vxs", _dp), ("vys", _dp), ("vzs", _dp), ("types", _ip), ("ids", _ip), ("engs", _fp), ("radii", _fp)]


class az_dump(C.Structure):
    _fields_ = [("n_atoms", C.c_int), ("n_cells", C.c_int), ("cnum", C.c_int * 3), ("csize", C.c_float * 3),
                ("crev", C.c_float * 3), ("cell_indexes", _ip), ("first_atom_in_cell", _ip), ("n_at_in_cell", _ip),
                ("ids", _ip), ("fx", _fp), ("fy", _fp), ("fz", _fp), ("x", _fp), ("y", _fp), ("z", _fp),
                ("vx", _fp), ("vy", _fp), ("vz", _fp)]


class az_engine_info_t(C.Structure):
    _fields_ = [("n_atoms", C.c_int), ("n_cells", C.c_int), ("cnum", C.c_int * 3), ("stencil", C.c_int * 3),
                ("csize", C.c_float * 3), ("crev", C.c_float * 3), ("pair_blocks", C.c_int), ("pair_threads", C.c_int),
                ("pair_smem_bytes", C.c_int), ("pair_zc", C.c_int), ("pair_vk", C.c_int), ("pair_ck", C.c_int),
                ("n_sm", C.c_int), ("max_neigh_seen", C.c_int), ("device_bytes", C.c_longlong),
                ("list_rows", C.c_int), ("flags", C.c_int), ("capacity", C.c_int), ("n_ranks", C.c_int),
                ("graph_replays", C.c_longlong), ("graphs_cached", C.c_int), ("peer_halo", C.c_int),
                ("tile_list", C.c_int), ("tile_atoms", C.c_int), ("tile_rows_smem", C.c_int),
                ("tile_atoms_seen", C.c_int), ("tile_runs_seen", C.c_int), ("reserved", C.c_int)]


# every symbol include/aztot_b200.h declares: name -> (restype, argtypes)
ENGINE_SYMBOLS = {
    "az_abi_version": (C.c_int, []),
    "az_last_error": (C.c_char_p, [C.c_void_p]),
    "az_device_count": (C.c_int, []),
    "az_engine_create": (C.c_int, [C.POINTER(az_system_desc), C.POINTER(az_cuda_opts), C.POINTER(C.c_void_p)]),
    "az_engine_destroy": (None, [C.c_void_p]),
    "az_engine_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "az_engine_sync": (C.c_int, [C.c_void_p]),
    "az_engine_stream": (C.c_void_p, [C.c_void_p]),
    "az_engine_stats": (C.c_int, [C.c_void_p, C.POINTER(az_stats)]),
    "az_engine_download": (C.c_int, [C.c_void_p, C.POINTER(az_atoms_out), C.c_int]),
    "az_engine_debug_dump": (C.c_int, [C.c_void_p, C.POINTER(az_dump)]),
    "az_engine_upload": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "az_engine_upload_async": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "az_engine_upload_f32_async": (C.c_int, [C.c_void_p, _fp, _fp, _fp, _fp, _fp, _fp]),
    "az_engine_download_f32": (C.c_int, [C.c_void_p, _fp, _fp, _fp, _fp, _fp, _fp]),
    "az_engine_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "az_engine_timers": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_longlong)]),
    "az_engine_launch_count": (C.c_longlong, [C.c_void_p]),
    "az_engine_step_timed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "az_host_alloc": (C.c_void_p, [C.c_size_t]),
    "az_host_free": (None, [C.c_void_p]),
    "az_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "az_engine_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "az_engine_info": (C.c_int, [C.c_void_p, C.POINTER(az_engine_info_t)]),
    "az_comm_unique_id": (C.c_int, [C.c_void_p]),
    "az_engine_rdf_begin": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "az_engine_rdf_sample": (C.c_int, [C.c_void_p]),
    "az_default_uvects": (C.c_int, [_dp, _dp, _dp]),
    "az_engine_rdf_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int, _ip, _ip, _ip]),
}
DECK_SYMBOLS = {
    "az_deck_load": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "az_deck_desc": (C.POINTER(az_system_desc), [C.c_void_p]),
    "az_deck_opts": (C.POINTER(az_cuda_opts), [C.c_void_p]),
    "az_deck_nsteps": (C.c_int, [C.c_void_p]),
    "az_deck_stat_period": (C.c_int, [C.c_void_p]),
    "az_deck_error": (C.c_char_p, [C.c_void_p]),
    "az_deck_warnings": (C.c_char_p, [C.c_void_p]),
    "az_deck_nstep_stat": (C.c_int, [C.c_void_p]),
    "az_deck_species_count": (C.c_int, [C.c_void_p, C.c_int]),
    "az_deck_rdf": (C.c_int, [C.c_void_p, _dp, _dp, _ip, _ip]),
    "az_deck_nuclei": (C.c_int, [C.c_void_p, _ip]),
    "az_deck_nucleus_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "az_deck_nucleus_count": (C.c_int, [C.c_void_p, C.c_int]),
    "az_deck_rdf_nuclei": (C.c_int, [C.c_void_p]),
    "az_photon_energy": (C.c_int, [C.c_double, C.c_double, C.c_double, _dp]),
    "az_deck_free": (None, [C.c_void_p]),
}


def bind(lib: C.CDLL, symbols=None, strict: bool = True) -> C.CDLL:
    """Attach restype/argtypes to every ABI symbol of ``lib``."""
    for table in ([ENGINE_SYMBOLS, DECK_SYMBOLS] if symbols is None else [symbols]):
        for name, (res, args) in table.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if strict:
                    raise
                continue
            fn.restype = res
            fn.argtypes = args
    return lib


STRICT_DECK = True
# AZTOT_B200_LIB: an A/B build of the same engine (profiles/build_variant.sh) for measurement jobs; never a different backend
LIB_PATH = os.environ.get("AZTOT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libaztot_b200.so")
_lib = None


def load_library() -> C.CDLL:
    """The nvcc-built engine.  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libaztot_b200.so is missing (%s): build it with `python -m aztotmd_b200.build` "
                "(needs nvcc; the engine is CUDA-only and has no CPU fallback)" % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH), strict=STRICT_DECK)
    return _lib
