"""Generate tests/golden/ref_serial_*.npz by running the reference's OWN serial build
(oracle/_ref/ref_serial_harness, compiled from /root/reference/src by oracle/build_ref.sh) on
small synthetic decks.  Run in the development container (the reference tree is not on the GPU
box); the fixtures are committed.  Each fixture holds the deck parameters, the positions exactly as
the reference parsed them, and the reference's double-precision forces and energies.

usage: python tests/golden/make_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from aztotmd_b200 import decks  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_serial_harness")


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            tag = f.read(8)
            if len(tag) < 8:
                break
            tag = tag.rstrip(b"\0").decode()
            (n,) = struct.unpack("i", f.read(4))
            e = np.frombuffer(f.read(32), dtype=np.float64)
            arrs = [np.frombuffer(f.read(8 * n), dtype=np.float64) for _ in range(9)]
            out[tag] = dict(engVdW=e[0], engCoul=e[1], engKin=e[2], engTot=e[3], pos=np.stack(arrs[0:3], 1),
                            vel=np.stack(arrs[3:6], 1), frc=np.stack(arrs[6:9], 1))
    return out


def run(deckdir, nsteps, mode, extra=None):
    dump = os.path.join(deckdir, "dump.bin")
    r = subprocess.run([HARNESS, str(nsteps), mode, dump, "1"] + ([extra] if extra else []), cwd=deckdir, stdin=subprocess.DEVNULL,
                       capture_output=True, text=True, check=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return read_dump(dump), line


def main():
    if not os.path.exists(HARNESS):
        subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        # LJ argon, all-pairs path of the serial twin (3 cells per axis: serial falls back to all_pair)
        d = os.path.join(tmp, "lj5")
        info = decks.lj_argon_deck(d, 5, nstep=2, init_vel="zero")
        cases.append(("lj_fcc5_allpairs", d, "full", dict(kind="lj", ncell=5, rc=info["rc"], box=[float(x) for x in info["box"]])))
        # LJ argon, linked-cell path of the serial twin (4 cells per axis)
        d = os.path.join(tmp, "lj6")
        info = decks.lj_argon_deck(d, 6, nstep=2, init_vel="zero")
        cases.append(("lj_fcc6_celllist", d, "fast", dict(kind="lj", ncell=6, rc=info["rc"], box=[float(x) for x in info["box"]])))
        # rock salt, Buckingham + damped-shifted-force Coulomb / direct Coulomb
        for elec in ("fenn", "dir"):
            d = os.path.join(tmp, "nacl_" + elec)
            info = decks.nacl_deck(d, 5, rc_vdw=9.0, r_real=9.0, alpha=0.2, elec=elec, init_vel="zero", nstep=2)
            cases.append(("nacl5_buck_" + elec, d, "full", dict(kind="nacl", ncell=5, elec=elec, rc_vdw=9.0, r_real=9.0,
                                                                alpha=0.2, box=[float(x) for x in info["box"]])))
        # rock salt, Buckingham + full Ewald sum (real space + reciprocal space + constant term): 'elec pme'
        d = os.path.join(tmp, "nacl_pme")
        info = decks.nacl_deck(d, 5, rc_vdw=9.0, r_real=9.0, alpha=0.3, elec="pme", init_vel="zero", nstep=2, jitter=0.4)
        cases.append(("nacl5_buck_pme", d, "full", dict(kind="nacl", ncell=5, elec="pme", rc_vdw=9.0, r_real=9.0, alpha=0.3,
                                                        ewald_k=[6, 6, 6], jitter=0.4, box=[float(x) for x in info["box"]])))
        # ... and an anisotropic box with different k ranges per axis: the k-space cut-off is the smallest of k_a / L_a
        d = os.path.join(tmp, "nacl_pme_aniso")
        info = decks.nacl_deck(d, (6, 5, 7), rc_vdw=9.0, r_real=9.0, alpha=0.3, elec="pme", init_vel="zero", nstep=2, jitter=0.4, ewald_k=(7, 5, 9))
        cases.append(("nacl657_buck_pme", d, "full", dict(kind="nacl", ncell=(6, 5, 7), elec="pme", rc_vdw=9.0, r_real=9.0, alpha=0.3,
                                                          ewald_k=[7, 5, 9], jitter=0.4, box=[float(x) for x in info["box"]])))
        # SURVEY 8f-4 groundwork: LJ argon with a list of constant bonds (harmonic and Morse), bonded terms added to the initial forces
        d = os.path.join(tmp, "lj5_bonds")
        info = decks.lj_argon_deck(d, 5, nstep=2, init_vel="zero")
        btypes = [(1, [5.0, 3.8, 0, 0, 0], "harm\t5.0\t3.8"), (2, [0.4, 1.5, 4.0, 0.4, 0], "mors\t0.4\t1.5\t4.0\t0.4")]
        with open(os.path.join(d, "field.txt"), "a") as ff:
            ff.write("bonds %d\n" % len(btypes) + "".join("%d\tAr\tAr\t%s\tcon\tcon\n" % (k + 1, t[2]) for k, t in enumerate(btypes)) + "bond_list 1\n")
        blist = [(4 * k, 4 * k + 1, 1 + (k % 2)) for k in range(0, info["n_atoms"] // 4, 3)] + [(2, 4 * 5 * 5 * 4 + 2, 1)]  # the last one spans a periodic face
        with open(os.path.join(d, "bonds.txt"), "w") as ff:
            ff.write("%d\n" % len(blist) + "".join("%d %d %d\n" % b for b in blist))
        cases.append(("lj5_bonds", d, "full", dict(kind="lj", ncell=5, rc=info["rc"], box=[float(x) for x in info["box"]], bonds=blist,
                                                   bond_types=[(t[0], t[1]) for t in btypes])))
        # ... and the same with valence angles (harmonic cosine) on top
        d = os.path.join(tmp, "lj5_bonds_angles")
        info = decks.lj_argon_deck(d, 5, nstep=2, init_vel="zero")
        atypes = [(6.0, -0.3), (2.5, 0.45)]
        with open(os.path.join(d, "field.txt"), "a") as ff:
            ff.write("bonds %d\n" % len(btypes) + "".join("%d\tAr\tAr\t%s\tcon\tcon\n" % (k + 1, t[2]) for k, t in enumerate(btypes)) + "bond_list 1\n")
            ff.write("angles %d\n" % len(atypes) + "".join("%d\tAr\thcos\t%g\t%g\n" % (k + 1, t[0], t[1]) for k, t in enumerate(atypes)) + "angle_list 1\n")
        with open(os.path.join(d, "bonds.txt"), "w") as ff:
            ff.write("%d\n" % len(blist) + "".join("%d %d %d\n" % b for b in blist))
        alist = [(4 * k, 4 * k + 1, 4 * k + 2, 1 + (k % 2)) for k in range(1, info["n_atoms"] // 4, 4)] + [(1, 3, 4 * 5 * 4 + 1, 2)]  # the last one reaches through a face
        with open(os.path.join(d, "angles.txt"), "w") as ff:
            ff.write("%d\n" % len(alist) + "".join("%d %d %d %d\n" % a for a in alist))
        cases.append(("lj5_bonds_angles", d, "full", dict(kind="lj", ncell=5, rc=info["rc"], box=[float(x) for x in info["box"]], bonds=blist,
                                                          bond_types=[(t[0], t[1]) for t in btypes], angles=alist, angle_types=atypes)))
        only = sys.argv[1:]
        for name, d, mode, meta in cases:
            if only and name not in only:
                continue
            dumps, line = run(d, 0, mode, "bonds" if "bonds" in meta else None)
            init = dumps["init"]
            names = [l.split()[0] for l in open(os.path.join(d, "atoms.xyz")).read().splitlines()[2:] if l.strip()]
            uniq = []
            for nm in names:
                if nm not in uniq:
                    uniq.append(nm)
            types = np.array([uniq.index(nm) for nm in names], dtype=np.int32)
            out = os.path.join(HERE, "ref_serial_%s.npz" % name)
            extra = {}
            import json
            if meta.get("elec") == "pme":
                j = json.loads(line)
                extra = dict(engElec1=j["engElec1"], engElec2=j["engElec2"])
            if "bonds" in meta:
                j = json.loads(line)
                assert j["n_bonds"] == len(meta["bonds"])
                extra = dict(engBond=j["engBond"], engAngle=j["engAngle"])
            np.savez_compressed(out, pos=init["pos"], frc=init["frc"], types=types, engVdW=init["engVdW"],
                                engCoul=init["engCoul"], meta=np.array(repr(meta)), harness=np.array(line), **extra)
            print(name, line)
            print("  ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
