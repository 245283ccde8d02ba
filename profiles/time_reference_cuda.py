"""Time the reference's own CUDA build (oracle/_ref/aztotmd_ref, see oracle/build_ref_cuda.sh) beside the new engine's
host driver on IDENTICAL decks on this box's GPU.  Step-loop time = (wall(n2) - wall(n1)) / (n2 - n1): two run
lengths cancel the init (the reference's O(nCell^2) init_bypass6 and its brute-force RDF pass at step 0).
usage: python profiles/time_reference_cuda.py [out.json]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aztotmd_b200 import decks  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "aztotmd_ref")
OURS = os.path.join(ROOT, "aztotmd_b200", "lib", "aztotmd")


def set_nstep(deck, n):
    p = os.path.join(deck, "control.txt")
    lines = open(p).read().splitlines()
    lines = [("nstep %d" % n) if l.startswith("nstep") else l for l in lines]
    open(p, "w").write("\n".join(lines) + "\n")


def wall(exe, deck, timeout=900):
    # line-buffered stdout: the reference build sometimes dies in its clean-up after "Finish." (exit code -11 on this
    # box); what matters for the timing is that the step loop completed, i.e. that the "Finish" line was printed
    t0 = time.perf_counter()
    r = subprocess.run(["stdbuf", "-oL", exe], cwd=deck, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=timeout)
    return time.perf_counter() - t0, r


def per_step(exe, deck, n1, n2):
    set_nstep(deck, n1)
    t1, r1 = wall(exe, deck)
    set_nstep(deck, n2)
    t2, r2 = wall(exe, deck)
    done = lambda r: r.returncode == 0 or "Finish" in (r.stdout or "")
    ok = done(r1) and done(r2)
    tail = (r2.stdout or "")[-200:]
    return dict(ok=ok, wall_n1=t1, wall_n2=t2, n1=n1, n2=n2, s_per_step=(t2 - t1) / (n2 - n1), rc=[r1.returncode, r2.returncode], tail=tail)


def main():
    out = {"what": "reference CUDA build (aztotmd_ref, sm_100a, unmodified arithmetic) vs new engine driver (aztotmd), same decks, same GPU",
           "cases": []}
    cases = [("fcc LJ argon 32 000 atoms, cell_list = rc = 8.488 A", lambda d: decks.lj_argon_deck(d, 20, nstep=10, stat=10 ** 6), 2000, 12000),
             ("fcc LJ argon 131 072 atoms, cell_list = rc = 8.488 A", lambda d: decks.lj_argon_deck(d, 32, nstep=10, stat=10 ** 6), 20, 80)]
    import test_deck
    import bench

    def cs2(d):
        s = bench.case_study_system("cs2")
        ctl = "\n".join(("rdf\t8.0   0.02\t1000000000\t1000000000\t-nucl" if l.startswith("rdf") else l) for l in test_deck.CS2_CONTROL.splitlines()) + "\nhist 1000000000\n"
        decks.write_deck(d, names="Ar", pos=s.pos, box=s.box, field_txt=test_deck.CS2_FIELD, control_txt=ctl,
                         cuda_txt="nstep stat 50\nnstep msdstat 50\nnthread a 16\nnthread b 32\n")
        return dict(n_atoms=len(s.pos))
    cases.append(("case-study-2-like: 4 000 abstract atoms in a 35 A box, surk + radiative thermostat, cell_list 2.7 (RDF off)", cs2, 2000, 12000))
    for name, make, n1, n2 in cases:
        with tempfile.TemporaryDirectory() as tmp:
            info = make(tmp)
            n = info["n_atoms"]
            row = {"deck": name, "n_atoms": n}
            for tag, exe in (("reference_cuda", REF), ("new_engine", OURS)):
                if not os.path.exists(exe):
                    row[tag] = {"ok": False, "why": "binary missing"}
                    continue
                try:
                    set_nstep(tmp, 5)
                    wall(exe, tmp)  # warm-up: pages the binary and the CUDA context in
                    r = per_step(exe, tmp, n1, n2)
                    r["atom_steps_per_s"] = n / r["s_per_step"] if r["s_per_step"] > 0 else None
                except subprocess.TimeoutExpired:
                    r = {"ok": False, "why": "timeout"}
                row[tag] = r
            a, b = row.get("reference_cuda", {}), row.get("new_engine", {})
            if a.get("ok") and b.get("ok") and a.get("atom_steps_per_s") and b.get("atom_steps_per_s"):
                row["speedup"] = b["atom_steps_per_s"] / a["atom_steps_per_s"]
            out["cases"].append(row)
            print(json.dumps(row), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
