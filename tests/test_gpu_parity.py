"""The parity tests proper: libaztot_b200.so (nvcc, sm_100a) through the C ABI on a B200 against
the CPU oracle on the same seeded inputs, plus size-independent properties at benchmark scale."""
import numpy as np
import pytest

import azt
import parity_cases
from aztotmd_b200 import abi
from aztotmd_b200.engine import Engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    lib = abi.load_library()
    assert lib.az_device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    return lib


@pytest.mark.parametrize("name,case", parity_cases.ALL_CASES, ids=[c[0] for c in parity_cases.ALL_CASES])
def test_cuda_matches_oracle(lib, name, case):
    case(lib)


def test_medium_lj_matches_oracle(lib):
    """32 000 atoms (20^3 fcc cells): the oracle still finishes in seconds."""
    s = azt.lj_system(ncell=(20, 20, 20), seed=21)
    o, e = parity_cases.run_parity(lib, s, nsteps=2)
    e.close()


def test_medium_ionic_matches_oracle(lib):
    s = azt.nacl_system(ncell=(10, 10, 10), r_real=12.0, rc_vdw=9.0, seed=22)  # 8 000 ions
    o, e = parity_cases.run_parity(lib, s, nsteps=2)
    e.close()


def test_case_study_2_size_surk_radiative(lib):
    s = azt.surk_system(n_side=16, spacing=1.3, box=35.0, seed=23, n_equil=4, eq_freq=2)  # 4096 atoms, droplet in vacuum
    o, e = parity_cases.run_parity(lib, s, nsteps=4, thermo=True, state_tol=2e-4, later_force_tol=1e-4)
    e.close()


@pytest.mark.parametrize("ncell,rebuild,skin,cell", [(100, 1, 0.0, None), (100, 10, 0.4, 8.9)], ids=["4M_rebuild1", "4M_rebuild10"])
def test_full_size_properties(lib, ncell, rebuild, skin, cell):
    """BASELINE config 3 (4 M atoms): properties that do not need the oracle --
    Newton's third law (sum of forces = 0 to rounding), bit-exact sortedness of the cell streams,
    persistent ids are a permutation, energy conservation over 50 steps, determinism."""
    s = azt.lj_system(ncell=(ncell,) * 3, seed=31, cell_size=cell)
    n = len(s.pos)
    e = Engine(s, lib=lib, rebuild_every=rebuild, skin=skin)
    e.step(0, 1)
    st0 = e.stats()
    d = e.dump()
    assert np.all(np.diff(d["cell"]) >= 0), "atoms are not sorted by cell"
    assert d["count"].sum() == n and d["first"][0] == 0 and np.array_equal(np.cumsum(d["count"])[:-1], d["first"][1:])
    assert np.array_equal(np.sort(d["ids"]), np.arange(n))
    # ascending persistent id inside each cell (canonical order)
    same = d["cell"][1:] == d["cell"][:-1]
    assert np.all(d["ids"][1:][same] > d["ids"][:-1][same])
    fsum = d["frc"].astype(np.float64).sum(0)
    assert np.abs(fsum).max() <= 1e-6 * np.abs(d["frc"]).astype(np.float64).sum()
    e.step(1, 49)
    st1 = e.stats()
    e0 = st0.engKin_d + st0.engVdW_d
    e1 = st1.engKin_d + st1.engVdW_d
    assert abs(e1 - e0) <= 2e-4 * abs(st0.engKin_d), (e0, e1)   # NVE drift over 50 fs relative to the kinetic energy
    e.close()


def test_no_silent_fallback(lib):
    """The loaded library is the nvcc build and really launches kernels."""
    s = azt.lj_system(ncell=(5, 5, 5))
    e = Engine(s, lib=lib)
    e.step(0, 2)
    assert e.launch_count >= 10
    e.close()
