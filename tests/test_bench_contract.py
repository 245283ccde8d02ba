"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON line with
the agreed keys (the reference's own serial build timed on the host cores), non-zero ranks of a torchrun launch of
that arm exit 0 without output, and the engine's own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_serial_harness")


def run_bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_serial_harness not built (oracle/build_ref.sh needs the reference tree)")
def test_reference_arm_prints_one_json_line_with_the_agreed_keys():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "atom-steps/s" and d["unit"] == "atom-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 1e4 and d["ms_per_step"] > 0
    assert "4,000,000 atoms" in d["config"]["workload"] and d["config"]["rebuild_every"] == 10
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] == d["value"] and "atom" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_needs_a_cuda_device():
    import ctypes as C
    from aztotmd_b200 import abi
    if abi.load_library().az_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--workload", "lj4k")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]  # and no number is printed
