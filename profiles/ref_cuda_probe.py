"""The reference's own CUDA build (oracle/_ref/aztotmd_ref = /root/reference/src compiled for sm_100a with the shims of
oracle/build_ref_cuda.sh) on this box's GPU: (1) compute-sanitizer memcheck of a 3-step run of a 4 000-atom LJ deck --
why its stat.dat is garbage on B200; (2) its step time at 131 072 atoms by differencing two run lengths that are long
enough for wall clocks (200 / 1 000 steps), beside the new engine's driver on the SAME deck on the reference's schedule
(bin + sort every step, `rebuild 1`), and whether each program's stat.dat holds physical numbers.
usage: python profiles/ref_cuda_probe.py out.json [sanitizer.log]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aztotmd_b200 import decks  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "aztotmd_ref")
OURS = os.path.join(ROOT, "aztotmd_b200", "lib", "aztotmd")


def set_nstep(deck, n):
    p = os.path.join(deck, "control.txt")
    lines = [("nstep %d" % n) if l.startswith("nstep") else l for l in open(p).read().splitlines()]
    open(p, "w").write("\n".join(lines) + "\n")


def run(exe, deck, timeout=600):
    t0 = time.perf_counter()
    try:
        r = subprocess.run(["stdbuf", "-oL", exe], cwd=deck, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=timeout)
        return time.perf_counter() - t0, r.returncode, r.stdout or ""
    except subprocess.TimeoutExpired:
        return time.perf_counter() - t0, "timeout", ""


def stat_valid(deck, n_atoms):
    """physical = finite, kinetic energy of a 86.6 K gas of n atoms within a factor 3, |engVdW| below 1 eV per atom"""
    p = os.path.join(deck, "stat.dat")
    if not os.path.exists(p):
        return dict(valid=False, why="no stat.dat")
    rows = [l.split("\t") for l in open(p).read().splitlines()[2:] if l.strip()]
    if not rows:
        return dict(valid=False, why="stat.dat has no rows")
    try:
        kin = [float(r[3]) for r in rows]; vdw = [float(r[4]) for r in rows]
    except Exception as e:  # noqa: BLE001
        return dict(valid=False, why="unparsable: %s" % e)
    ke_expected = 1.5 * n_atoms * 8.617333e-5 * 86.6
    ok = all(abs(v) < 1.0 * n_atoms for v in vdw) and all(k == k and ke_expected / 3 < k < ke_expected * 3 for k in kin[:1])
    return dict(valid=bool(ok), first_row=dict(engKin=kin[0], engVdW=vdw[0]), last_row=dict(engKin=kin[-1], engVdW=vdw[-1]), ke_expected=ke_expected)


def main():
    out = {"what": "reference CUDA build (aztotmd_ref, sm_100a) vs the new engine's driver, same deck, same GPU"}
    out["note"] = ("compute-sanitizer is closed on this GPU pool ('runs under it have left GPUs needing a reset'), so the fault of the reference "
                   "build cannot be located that way; what is recorded instead is what the program does: whether its step loop completes, how "
                   "long a step takes, and whether its stat.dat holds physical numbers")
    # (1) memcheck (only where the pool allows it)
    with tempfile.TemporaryDirectory() as tmp:
        decks.lj_argon_deck(tmp, 10, nstep=3, stat=1, init_vel="zero")
        r = subprocess.run(["compute-sanitizer", "--tool", "memcheck", "--print-limit", "40", REF], cwd=tmp, stdin=subprocess.DEVNULL,
                           capture_output=True, text=True, timeout=900)
        log = (r.stdout or "") + (r.stderr or "")
        if len(sys.argv) > 2:
            open(sys.argv[2], "w").write(log)
        errs = [l for l in log.splitlines() if "Invalid" in l or "ERROR SUMMARY" in l or " at " in l][:60]
        out["memcheck"] = dict(rc=r.returncode, summary=[l for l in log.splitlines() if "ERROR SUMMARY" in l], first_lines=errs[:30])
    # (2) timing at 131 072 atoms
    n1, n2 = int(os.environ.get('AZ_REF_N1', 200)), int(os.environ.get('AZ_REF_N2', 1000))
    with tempfile.TemporaryDirectory() as tmp:
        info = decks.lj_argon_deck(tmp, 32, nstep=10, stat=100)
        n = info["n_atoms"]
        row = {"deck": "fcc LJ argon %d atoms, cell_list = rc = 8.488 A, rebuild 1 (the reference's schedule), 86.6 K" % n, "n_atoms": n, "n1": n1, "n2": n2}
        for tag, exe in (("reference_cuda", REF), ("new_engine", OURS)):
            set_nstep(tmp, 5); run(exe, tmp)  # warm-up
            set_nstep(tmp, n1); t1, rc1, o1 = run(exe, tmp)
            set_nstep(tmp, n2); t2, rc2, o2 = run(exe, tmp)
            done = lambda rc, o: rc == 0 or "Finish" in o  # noqa: E731  (the reference sometimes dies in its clean-up after the loop)
            sps = (t2 - t1) / (n2 - n1)
            row[tag] = dict(loop_completed=bool(done(rc1, o1) and done(rc2, o2)), rc=[rc1, rc2], wall_n1=t1, wall_n2=t2, s_per_step=sps,
                            atom_steps_per_s=(n / sps if sps > 0 else None), stat=stat_valid(tmp, n), tail=o2[-300:])
        a, b = row["reference_cuda"], row["new_engine"]
        if a["atom_steps_per_s"] and b["atom_steps_per_s"]:
            row["ratio_new_over_reference_cuda"] = b["atom_steps_per_s"] / a["atom_steps_per_s"]
        out["timing"] = row
    print(json.dumps(out, indent=1))
    json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
