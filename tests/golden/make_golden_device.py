"""Generate tests/golden/ref_device_vectors.npz: outputs of the reference's OWN __device__ functions (pair functions of
src/cuVdW.cu and src/cuElec.cu, get_angled_vector / adsorb_rand_photon / radiate_photon3 of src/cuTemp.cu) executed
on the CPU by oracle/_ref/libref_device.so (oracle/build_ref_device.sh; needs /root/reference, i.e. this container).
The vectors pin the arithmetic that exists only in the reference's CUDA build.  Run:  python tests/golden/make_golden_device.py"""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import azt  # noqa: E402
from aztotmd_b200.engine import KB, M_SCALE  # noqa: E402

_fp = C.POINTER(C.c_float)


def main():
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref_device.sh")], check=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_device.so"))
    lib.refdev_vdw.argtypes = [C.c_int, _fp, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, _fp, _fp]
    lib.refdev_coul.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, _fp]
    lib.refdev_angled_vector.argtypes = [_fp, C.c_float, C.c_float, _fp]
    lib.refdev_tstat_atom.argtypes = [_fp, _fp, C.c_float, C.c_float, C.POINTER(C.c_uint), _fp, C.c_float, C.c_float, C.c_float, _fp]
    rng = np.random.default_rng(20261018)
    out = {}

    # ---- van der Waals forms: kind, prepared parameters p0..p4, r^2, radii -> F/r, U (with and without a known r)
    params = {1: [4 * 0.01006, 3.3952 ** 2, 24 * 0.01006, 0, 0], 2: [1256.31, 0.317, 7.0, 0, 0], 4: [0.25, 3.0, 3.2, 70.0, 140.0],
              5: [900.0, 0.33, 0.02, 0, 0], 6: [900.0, 0.33, 1.5, 0, 0], 7: [75.0, 8.0, 1.0, 1.0, 0]}
    rows = []
    for kind, p in params.items():
        pp = (C.c_float * 5)(*p)
        for _ in range(48):
            r = float(np.float32(rng.uniform(2.2, 9.0) if kind != 7 else rng.uniform(1.0, 6.0)))
            r2 = float(np.float32(r * r))
            rad = [float(np.float32(rng.uniform(0.577, 0.603))) for _ in range(2)]
            for with_r in (0, 1):
                f, e = C.c_float(), C.c_float()
                rin = float(np.float32(math.sqrt(r2))) if with_r else 0.0
                assert lib.refdev_vdw(kind, pp, r2, rin, with_r, rad[0], rad[1], C.byref(f), C.byref(e)) == 0
                rows.append([kind, with_r, r2, rad[0], rad[1], f.value, e.value] + list(p))
    out["vdw"] = np.array(rows, dtype=np.float64)

    # ---- real-space Coulomb: kind, r^2, q_i q_j -> F/r, U, r
    alpha, rc = 0.2, 12.0
    daipi2 = 2 * alpha / math.sqrt(3.14159265359)
    c1 = math.erfc(alpha * rc) / rc
    c2 = math.erfc(alpha * rc) / rc ** 2 + daipi2 * math.exp(-(alpha * rc) ** 2) / rc
    rows = []
    for kind in (1, 2, 3):
        for _ in range(48):
            r = float(np.float32(rng.uniform(2.0, 12.0)))
            r2 = float(np.float32(r * r))
            q = float(rng.choice([-1.0, 1.0, 0.64, -2.0]))
            f, e, ro = C.c_float(), C.c_float(), C.c_float()
            assert lib.refdev_coul(kind, r2, q, alpha, daipi2, c1, c2, rc, C.byref(f), C.byref(e), C.byref(ro)) == 0
            rows.append([kind, r2, q, f.value, e.value, ro.value])
    out["coul"] = np.array(rows, dtype=np.float64)
    out["coul_consts"] = np.array([alpha, daipi2, c1, c2, rc], dtype=np.float64)

    # ---- get_angled_vector
    rows = []
    for _ in range(64):
        v = rng.standard_normal(3).astype(np.float32)
        if rng.random() < 0.15:
            v[0] = 0.0  # the v1.x == 0 branches
        cos_phi = float(np.float32(rng.uniform(-1.0, 1.0)))
        theta = float(np.float32(rng.uniform(0.0, 6.28)))
        o = (C.c_float * 3)()
        lib.refdev_angled_vector(v.ctypes.data_as(_fp), cos_phi, theta, o)
        rows.append(list(v) + [cos_phi, theta] + list(o))
    out["angled"] = np.array(rows, dtype=np.float64)

    # ---- one atom through the thermostat body
    ux, uy, uz = (np.zeros(3072) for _ in range(3))
    ol = azt.oracle_lib()
    dp = C.POINTER(C.c_double)
    ol.azo_default_uvects(ux.ctypes.data_as(dp), uy.ctypes.data_as(dp), uz.ctypes.data_as(dp))
    uv = np.stack([ux, uy, uz], 1).astype(np.float32).copy()
    mass = float(np.float32(39.9 * M_SCALE))
    rows = []
    while len(rows) < 96:
        v = (rng.standard_normal(3) * 3.0).astype(np.float32)
        U = float(np.float32(rng.choice([0.0, rng.uniform(0, 0.05), rng.uniform(0.05, 0.3)])))
        pe = float(np.float32(rng.gamma(5.0, KB * 500.0)))
        rnd = rng.integers(0, 2 ** 31, size=3).astype(np.uint32)
        # the reference leaves the radiated direction uninitialised when e/(m c v0) >= 1 (src/cuTemp.cu:663-675): only
        # vectors on the defined branch are kept
        vv, UU = v.copy(), np.float32(U)
        u = uv[int(rnd[0]) % 3072]
        ermc = np.float32(pe) * np.float32(3.33567e-5) / np.float32(mass)
        v1 = vv + ermc * u
        U1 = UU + np.float32(pe) + np.float32(0.5) * np.float32(mass) * (np.float32((vv * vv).sum()) - np.float32((v1 * v1).sum()))
        if U1 > 1e-4:
            a = np.float32(0.9) * U1 * np.float32(3.33567e-5) / np.float32(mass)
            if not a / np.sqrt((v1 * v1).sum()) < 0.98:
                continue
        vel = v.copy()
        Uc = C.c_float(U)
        rad = C.c_float()
        lib.refdev_tstat_atom(vel.ctypes.data_as(_fp), C.byref(Uc), mass, pe, rnd.ctypes.data_as(C.POINTER(C.c_uint)),
                              uv.ctypes.data_as(_fp), 2.73, 4.731, 0.2, C.byref(rad))
        rows.append(list(v) + [U, pe, float(rnd[0]), float(rnd[1]), float(rnd[2])] + list(vel) + [Uc.value, rad.value])
    out["tstat"] = np.array(rows, dtype=np.float64)

    # ---- grid geometry (split_cells), cell-pair existence per axis (cell_dist_int), wrap + bin (put_periodic, count_cell)
    _ip = C.POINTER(C.c_int)
    lib.refdev_split_cells.argtypes = [_fp, C.c_float, C.c_int, _ip, _fp, _fp]
    lib.refdev_cell_dist.restype = C.c_int
    lib.refdev_cell_dist.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp, _fp, _ip]
    lib.refdev_wrap_and_bin.argtypes = [_fp, _fp, C.c_float, _fp, _ip, _fp, _ip, _ip, _fp]
    rows = []
    boxes = [(1141.5, 1141.5, 1141.5, 85.0), (35.0, 35.0, 35.0, 2.7), (570.26, 570.26, 570.26, 8.488), (570.26, 570.26, 570.26, 8.738)]
    for _ in range(60):
        boxes.append((float(np.float32(rng.uniform(8, 300))), float(np.float32(rng.uniform(8, 300))), float(np.float32(rng.uniform(8, 300))),
                      float(np.float32(rng.uniform(2.0, 14.0)))))
    for lx, ly, lz, r in boxes:
        if min(lx, ly, lz) < r:
            continue
        L = (C.c_float * 3)(lx, ly, lz)
        cn, cs, cr = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        lib.refdev_split_cells(L, r, 1000, cn, cs, cr)
        rows.append([L[0], L[1], L[2], np.float32(r)] + list(cn) + list(cs) + list(cr))
    out["grid"] = np.array(rows, dtype=np.float64)
    rows = []
    for mx in range(1, 10):
        for csize, rskip in ((3.0, 2.9), (3.0, 3.1), (2.9, 6.0), (2.917, 8.0), (8.8, 8.488)):
            for xi in range(mx):
                for xj in range(mx):
                    rmin, rmax, sh = C.c_float(), C.c_float(), C.c_int()
                    skip = lib.refdev_cell_dist(xi, xj, mx, csize, rskip, C.byref(rmin), C.byref(rmax), C.byref(sh))
                    rows.append([mx, csize, rskip, xi, xj, skip, sh.value, rmin.value])
    out["celldist"] = np.array(rows, dtype=np.float64)
    rows = []
    for _ in range(400):
        lx, ly, lz, r = boxes[int(rng.integers(0, len(boxes)))]
        if min(lx, ly, lz) < r:
            continue
        L = (C.c_float * 3)(lx, ly, lz)
        cn, cs, cr = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        lib.refdev_split_cells(L, r, 1000, cn, cs, cr)
        mode = rng.integers(0, 4)
        x = np.array([rng.uniform(0, lx), rng.uniform(0, ly), rng.uniform(0, lz)], dtype=np.float32)
        if mode == 1:   # just outside a face
            a = int(rng.integers(0, 3)); x[a] = np.float32([lx, ly, lz][a] + rng.uniform(0, 0.5)) if rng.random() < 0.5 else np.float32(-rng.uniform(0, 0.5))
        elif mode == 2:  # exactly on a cell or box boundary
            a = int(rng.integers(0, 3)); x[a] = np.float32(cs[a] * int(rng.integers(0, cn[a] + 1)))
        elif mode == 3:  # one ulp below the upper face
            a = int(rng.integers(0, 3)); x[a] = np.nextafter(np.float32([lx, ly, lz][a]), np.float32(0))
        v = rng.standard_normal(3).astype(np.float32)
        xin = x.copy()
        cell, cross, mom = C.c_int(), (C.c_int * 6)(), (C.c_float * 6)()
        lib.refdev_wrap_and_bin(x.ctypes.data_as(_fp), v.ctypes.data_as(_fp), np.float32(0.004135), L, cn, cr, C.byref(cell), cross, mom)
        rows.append(list(xin) + list(v) + [L[0], L[1], L[2]] + list(cn) + list(cr) + list(x) + [cell.value] + list(cross) + list(mom))
    out["wrapbin"] = np.array(rows, dtype=np.float64)

    # ---- Nose-Hoover scalar recurrence: 40 consecutive blocks from three starting points
    lib.refdev_nose.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, _fp, _fp]
    rows = []
    for teKin, tau, ke0 in ((19.4, 0.02, 48.0), (5.0, 0.1, 1.0), (120.0, 0.5, 119.0)):
        rq, qt = np.float32(0.5 / teKin / tau / tau), np.float32(2.0 * teKin)
        ke, chit, ci, ts = C.c_float(ke0), C.c_float(0.0), C.c_float(0.0), C.c_float(0.0)
        for k in range(40):
            kin = ke.value
            lib.refdev_nose(np.float32(0.001), np.float32(teKin), rq, qt, C.byref(ke), C.byref(chit), C.byref(ci), C.byref(ts))
            rows.append([0.001, teKin, float(rq), float(qt), kin, ke.value, chit.value, ci.value, ts.value])
            ke = C.c_float(ke.value * (1.0 + 0.01 * math.sin(k)))  # the integrator changes the kinetic energy between blocks
    out["nose"] = np.array(rows, dtype=np.float64)
    out["tstat_mass"] = np.array([mass])
    if "bonds" not in sys.argv[1:]:  # the committed file is only rewritten on request: `make_golden_device.py all`
        if "all" in sys.argv[1:] or not os.path.exists(os.path.join(HERE, "ref_device_vectors.npz")):
            np.savez_compressed(os.path.join(HERE, "ref_device_vectors.npz"), **out)
    bonds_and_angles(lib)


def bonds_and_angles(lib):
    """SURVEY 8f-4 groundwork: the reference's bond potentials (src/cuBonds.cu:1123-1224) and the harmonic-cosine valence
    angle (src/cuAngles.cu:230-283, with delta_periodic) -> tests/golden/ref_device_bonds_angles.npz"""
    lib.refdev_bond.argtypes = [C.c_int, _fp, C.c_float, C.c_float, _fp, _fp]
    lib.refdev_angle_hcos.argtypes = [C.c_float, C.c_float, _fp, _fp, _fp, _fp]
    rng = np.random.default_rng(8642)
    params = {1: [25.0, 1.1, 0, 0, 0], 2: [4.5, 2.0, 0.96, 4.5, 0], 3: [0.3, 2.0, 2.1, 0.3, 1.0], 4: [1256.31, 0.317, 7.0, 0, 0],
              5: [900.0, 0.33, 1.5, 3.0, 5.0]}
    rows = []
    for kind, p in params.items():
        pp = (C.c_float * 5)(*p)
        for _ in range(40):
            r = float(np.float32(rng.uniform(0.8, 3.0)))
            r2 = float(np.float32(r * r))
            f, e = C.c_float(), C.c_float()
            assert lib.refdev_bond(kind, pp, r2, r, C.byref(f), C.byref(e)) == 0
            rows.append([kind, r2, r, f.value, e.value] + list(p))
    out = {"bond": np.array(rows, dtype=np.float64)}
    rows = []
    L = np.array([20.0, 24.0, 28.0], dtype=np.float32)
    for _ in range(120):
        k, cos0 = float(np.float32(rng.uniform(1.0, 8.0))), float(np.float32(math.cos(rng.uniform(1.6, 2.2))))
        c = rng.uniform(0.0, 1.0, 3) * L
        l1 = c + rng.normal(0.0, 1.0, 3)
        l2 = c + rng.normal(0.0, 1.0, 3)
        xyz = (np.stack([c, l1, l2]) % L).astype(np.float32)  # ligands may sit across a periodic face
        frc = np.zeros(9, dtype=np.float32)
        e = C.c_float()
        lib.refdev_angle_hcos(k, cos0, xyz.ravel().ctypes.data_as(_fp), L.ctypes.data_as(_fp), frc.ctypes.data_as(_fp), C.byref(e))
        rows.append([k, cos0] + [float(v) for v in xyz.ravel()] + [float(v) for v in frc] + [e.value])
    out["angle"] = np.array(rows, dtype=np.float64)
    out["angle_box"] = L.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_device_bonds_angles.npz"), **out)
    print("bonds/angles:", len(out["bond"]), "bond vectors,", len(out["angle"]), "angle vectors")
    print("wrote ref_device_vectors.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
