"""Pin the engine against the reference's own CUDA build (oracle/_ref/aztotmd_ref) on this box's GPU through the one
numeric surface that build offers: stat.dat.  Decks start from zero velocities, so every row is deterministic
(no time-seeded rand() in the reference's velocity init); both programs read the SAME four files.
  * LJ argon: engKin / engVdW / engTot of the first steps must agree to rounding (1e-5): pins binning, cell-pair
    traversal, image shifts, the LJ pair function and both Verlet stages of the real CUDA build.
  * surk + radiative thermostat (case-study-2-like): step 0 only (the reference's thermostat RNG is unseeded and its
    initial radii are 0.577 + rand*1e-4): engVdW must agree to 2e-3: pins the radius-dependent pair function.
usage: python profiles/compare_reference_cuda_stats.py [out.json]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aztotmd_b200 import decks  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "aztotmd_ref")
OURS = os.path.join(ROOT, "aztotmd_b200", "lib", "aztotmd")


def run(exe, deck):
    r = subprocess.run(["stdbuf", "-oL", exe], cwd=deck, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600)
    rows = []
    p = os.path.join(deck, "stat.dat")
    if os.path.exists(p):
        lines = open(p).read().splitlines()
        hdr = lines[0].split("\t")
        for l in lines[2:]:
            v = l.split("\t")
            if len(v) >= len(hdr):
                rows.append({k: float(x) for k, x in zip(hdr, v)})
        os.rename(p, p + "." + os.path.basename(exe))
    return r.returncode, rows, (r.stdout or "")[-300:]


def main():
    import bench
    import test_deck
    out = {"cases": []}

    def lj(d):
        return decks.lj_argon_deck(d, 12, nstep=6, stat=1, init_vel="zero", jitter=0.3, seed=3)

    def cs2(d):
        s = bench.case_study_system("cs2")
        ctl = "\n".join(("rdf\t8.0   0.02\t1000000000\t1000000000\t-nucl" if l.startswith("rdf") else
                         ("nstep 2" if l.startswith("nstep") else ("stat\t\t1" if l.startswith("stat") else l)))
                        for l in test_deck.CS2_CONTROL.splitlines()) + "\nhist 1000000000\n"
        decks.write_deck(d, names="Ar", pos=s.pos, box=s.box, field_txt=test_deck.CS2_FIELD, control_txt=ctl,
                         cuda_txt="nstep stat 1\nnstep msdstat 1\nnthread a 16\nnthread b 32\n")
        return dict(n_atoms=len(s.pos))

    for name, make, keys, tol, nrows in (("fcc LJ argon 6 912 atoms, jitter 0.3 A, zero initial velocities", lj, ("engKin", "engVdW", "engTot"), 1e-5, 5),
                                         ("case-study-2-like 4 000 atoms, surk + radiative thermostat", cs2, ("engVdW",), 2e-3, 1)):
        with tempfile.TemporaryDirectory() as tmp:
            make(tmp)
            # the reference flushes stat.dat every 'nstep stat' samples
            cu = os.path.join(tmp, "cuda.txt")
            open(cu, "w").write("nstep stat 1\nnstep msdstat 1\nnthread a 16\nnthread b 32\n")
            rc_r, rows_r, tail_r = run(REF, tmp)
            rc_o, rows_o, tail_o = run(OURS, tmp)
            row = {"deck": name, "reference_rc": rc_r, "engine_rc": rc_o, "reference_rows": rows_r[:nrows], "engine_rows": rows_o[:nrows], "tolerance": tol}
            worst = 0.0
            ok = len(rows_r) >= nrows and len(rows_o) >= nrows
            if ok:
                for a, b in zip(rows_r[:nrows], rows_o[:nrows]):
                    scale = max(abs(a.get("engVdW", 0.0)), abs(a.get("engKin", 0.0)), 1e-12)
                    for k in keys:
                        worst = max(worst, abs(a[k] - b[k]) / scale)
            row["worst_relative_difference"] = worst if ok else None
            row["agree"] = bool(ok and worst <= tol)
            if not ok:
                row["reference_tail"] = tail_r
                row["engine_tail"] = tail_o
            out["cases"].append(row)
            print(json.dumps(row), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
