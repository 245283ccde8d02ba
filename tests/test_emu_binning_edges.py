"""Cell assignment on awkward inputs, engine (emulator build of the kernels) against the oracle, BIT-EXACT: random
rectangular boxes and cell sizes, atoms exactly on cell boundaries and box faces, one ulp inside / outside them, and
outside the box by up to two box lengths (put_periodic wraps before count_cell bins, src/cuMDfunc.cu:35-106,
src/cuSort.cu:114-128).  The oracle itself is pinned on the reference's device source for these rules
(tests/golden/ref_device_vectors.npz: 'wrapbin', 'grid')."""
import numpy as np
import pytest

import azt
import parity_cases
from aztotmd_b200.engine import Engine, Species, System


def awkward_system(seed):
    rng = np.random.default_rng(seed)
    box = np.array([float(np.float32(rng.uniform(18.0, 60.0))) for _ in range(3)])
    cell = float(np.float32(rng.uniform(3.0, 9.0)))
    n = 160
    pos = rng.uniform(0.0, 1.0, (n, 3)) * box
    ncell = np.floor(box.astype(np.float32) / np.float32(cell)).astype(int)
    csize = box / ncell
    k = 0
    for a in range(3):                       # atoms exactly on / one ulp around cell boundaries and faces of every axis
        for m in (0, 1, ncell[a] // 2, ncell[a] - 1, ncell[a]):
            x = np.float32(m * csize[a])
            for v in (x, np.nextafter(x, np.float32(-1e9)), np.nextafter(x, np.float32(1e9))):
                pos[k, a] = float(v)
                k += 1
    pos[k:k + 12] += rng.integers(-2, 3, (12, 3)) * box   # outside the box: wrapped by whole box lengths first
    s = System(pos=pos, types=np.zeros(n, dtype=np.int32), box=box, species=[Species("Ar", 39.9)], timestep=0.001, cell_size=cell)
    s.set_vdw(0, 0, "lnjs", min(cell, 3.0), [1e-6, 1.0])    # a force field that stays tiny whatever the distances are
    return s


@pytest.mark.parametrize("seed", range(16))
def test_cells_of_awkward_positions_are_bit_exact(seed):
    s = awkward_system(seed)
    o = azt.Oracle(s)
    e = Engine(s, lib=azt.emu_lib())
    try:
        o.step(0, 1); e.step(0, 1)
        do, de = o.dump(), e.dump()
        parity_cases.assert_cells_equal(do, de)
        assert np.array_equal(do["pos"], de["pos"])          # wrapped coordinates, bit for bit
        assert do["count"].sum() == len(s.pos)
    finally:
        e.close()
