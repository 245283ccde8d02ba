"""Synthetic input decks in the reference's own file formats.

The generators write real ``atoms.xyz`` / ``field.txt`` / ``control.txt`` / ``cuda.txt`` files
(formats: /root/reference/src/sys_init.cpp:174-565,590-988, src/cuInit.cu:684-754) so the same
files feed the reference's serial build, the oracle and the CUDA engine.  Configurations follow
SURVEY.md section 8(d) / BASELINE.md 4.3.
"""
from __future__ import annotations

import os
import numpy as np

# LJ argon of case study 1 (case study 1/field.txt:5)
AR_EPS = 0.01006
AR_SIGMA = 3.3952
AR_MASS = 39.9


def fcc_positions(ncell, a, jitter=0.05, seed=1):
    """ncell = (nx,ny,nz) fcc unit cells of edge a; seeded jitter avoids exact cell-boundary ties."""
    nx, ny, nz = ncell
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    pos = (cells[:, None, :] + basis[None, :, :] + 0.25).reshape(-1, 3) * a
    if jitter:
        rng = np.random.default_rng(seed)
        pos += rng.uniform(-jitter, jitter, size=pos.shape)
    box = np.array([nx * a, ny * a, nz * a])
    pos %= box
    return pos, box


def rocksalt_positions(ncell, a, jitter=0.05, seed=2):
    """ncell^3 rock-salt unit cells (8 ions each); returns positions, type (0 = Na, 1 = Cl), box."""
    nx, ny, nz = ncell
    na = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    cl = na + np.array([0.5, 0, 0])
    basis = np.concatenate([na, cl]) % 1.0
    btype = np.array([0] * 4 + [1] * 4)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    pos = (cells[:, None, :] + basis[None, :, :] + 0.125).reshape(-1, 3) * a
    types = np.tile(btype, len(cells))
    if jitter:
        rng = np.random.default_rng(seed)
        pos += rng.uniform(-jitter, jitter, size=pos.shape)
    box = np.array([nx * a, ny * a, nz * a])
    pos %= box
    return pos, types, box


def write_atoms_xyz(path, names, pos, box):
    """atoms.xyz: N / '1 la lb lc' / 'name x y z' lines (src/sys_init.cpp:487-565)."""
    with open(path, "w") as f:
        f.write("%d\n" % len(pos))
        f.write("1 %f %f %f\n" % tuple(box))
        if isinstance(names, str):
            lines = ["%s\t%f\t%f\t%f" % (names, p[0], p[1], p[2]) for p in pos]
        else:
            lines = ["%s\t%f\t%f\t%f" % (nm, p[0], p[1], p[2]) for nm, p in zip(names, pos)]
        f.write("\n".join(lines))
        f.write("\n")


def write_deck(dirname, *, names, pos, box, field_txt, control_txt, cuda_txt=None):
    os.makedirs(dirname, exist_ok=True)
    write_atoms_xyz(os.path.join(dirname, "atoms.xyz"), names, pos, box)
    with open(os.path.join(dirname, "field.txt"), "w") as f:
        f.write(field_txt)
    with open(os.path.join(dirname, "control.txt"), "w") as f:
        f.write(control_txt)
    with open(os.path.join(dirname, "cuda.txt"), "w") as f:
        f.write(cuda_txt or "nstep stat 50\nnstep msdstat 50\nnthread a 16\nnthread b 32\n")
    return dirname


def lj_argon_deck(dirname, ncell, *, rho_star=0.8442, rc_sigma=2.5, temperature=86.6,
                  nstep=100, init_vel="gaus", tstat="none", cell_list=None, jitter=0.05, seed=1,
                  rebuild_every=1, skin=0.0, stat=1000, rdf_freq=10 ** 9, rdf_rmax=8.0, rdf_dr=0.02):
    """fcc LJ argon (config 3/5 of BASELINE.json): a = (4/rho)^(1/3) sigma, rc = 2.5 sigma."""
    if isinstance(ncell, int):
        ncell = (ncell, ncell, ncell)
    a = (4.0 / rho_star) ** (1.0 / 3.0) * AR_SIGMA
    rc = rc_sigma * AR_SIGMA
    pos, box = fcc_positions(ncell, a, jitter=jitter, seed=seed)
    field = "spec 1\nAr\tAr\t%g\t0.0\t0.0\nred-ox 0\nvdw 1\nAr\tAr\tlnjs\t%.6f\t%g\t%g\n" % (
        AR_MASS, rc, AR_EPS, AR_SIGMA)
    ctl = (
        "timestep 0.001 ps\nnstep %d\nnequil 0\neqfreq 100\ntemperature %g\t%s\n"
        "init_vel\t%s\npermittivity 1.0\ncell_list\t%.6f\nmax_neigh\t185\nelec\tnone\n"
        "rdf\t%g %g\t%d\t%d\nhist %d\nstat\t%d\nseed %d\n"
        % (nstep, temperature, tstat, init_vel, cell_list if cell_list else rc,
           rdf_rmax, rdf_dr, rdf_freq, 10 ** 9, 10 ** 9, stat, seed))
    cuda = "nstep stat 50\nnstep msdstat 50\nnthread a 16\nnthread b 32\nrebuild %d\nskin %g\n" % (
        rebuild_every, skin)
    write_deck(dirname, names="Ar", pos=pos, box=box, field_txt=field, control_txt=ctl, cuda_txt=cuda)
    return dict(n_atoms=len(pos), box=box, rc=rc, a=a)


def nacl_deck(dirname, ncell, *, a=5.9, rc_vdw=9.0, r_real=12.0, alpha=0.2, temperature=1200.0,
              nstep=100, elec="fenn", init_vel="gaus", jitter=0.05, seed=2, cell_list=None, stat=1000, ewald_k=(6, 6, 6)):
    """Rock-salt NaCl-like melt (config 4): Buckingham (Tosi-Fumi-like) + real-space Coulomb."""
    if isinstance(ncell, int):
        ncell = (ncell, ncell, ncell)
    pos, types, box = rocksalt_positions(ncell, a, jitter=jitter, seed=seed)
    names = np.where(types == 0, "Na", "Cl")
    field = (
        "spec 2\nNa\tNa\t22.99\t1.0\t0.0\nCl\tCl\t35.45\t-1.0\t0.0\nred-ox 0\nvdw 3\n"
        "Na\tNa\tbuck\t%g\t424.097\t0.317\t1.05\n"
        "Na\tCl\tbuck\t%g\t1256.31\t0.317\t7.00\n"
        "Cl\tCl\tbuck\t%g\t3488.998\t0.317\t72.40\n" % (rc_vdw, rc_vdw, rc_vdw))
    if elec == "fenn":
        el = "elec\tfenn\t%g\t%g\n" % (r_real, alpha)
    elif elec == "dir":
        el = "elec\tdir\t%g\n" % r_real
    elif elec == "pme":
        el = "elec\tpme\t%g\t%g\t%d\t%d\t%d\n" % (r_real, alpha, ewald_k[0], ewald_k[1], ewald_k[2])
    else:
        el = "elec\tnone\n"
    ctl = (
        "timestep 0.001 ps\nnstep %d\nnequil 0\neqfreq 100\ntemperature %g\tnone\n"
        "init_vel\t%s\npermittivity 1.0\ncell_list\t%.6f\nmax_neigh\t185\n%s"
        "rdf\t8.0 0.02\t%d\t%d\nhist %d\nstat\t%d\nseed %d\n"
        % (nstep, temperature, init_vel, cell_list if cell_list else max(r_real, rc_vdw), el,
           10 ** 9, 10 ** 9, 10 ** 9, stat, seed))
    write_deck(dirname, names=names, pos=pos, box=box, field_txt=field, control_txt=ctl)
    return dict(n_atoms=len(pos), box=box)
