"""The C ABI (include/aztot_b200.h) against its ctypes mirror and the built library, without a GPU:
every function the header declares is exported by libaztot_b200.so and bound in aztotmd_b200/abi.py, every struct has
the size and field offsets the C compiler gives it, and the library refuses to work without a device instead of
falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from aztotmd_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aztot_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|char|long long|az_\w+)\s*\*?\s*\*?\s*(az_\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_every_declared_function_is_bound_and_exported():
    names = declared_functions()
    assert len(names) >= 30 and "az_engine_create" in names and "az_comm_unique_id" in names and "az_deck_load" in names
    bound = set(abi.ENGINE_SYMBOLS) | set(abi.DECK_SYMBOLS)
    assert set(names) == bound, (sorted(set(names) - bound), sorted(bound - set(names)))
    lib = C.CDLL(abi.LIB_PATH)  # the nvcc build; loading needs no GPU
    for n in names:
        assert hasattr(lib, n), "libaztot_b200.so does not export %s" % n


def test_struct_layouts_match_the_c_compiler(tmp_path):
    structs = ["az_species", "az_vdw_entry", "az_bond_type", "az_angle_type", "az_system_desc", "az_cuda_opts", "az_stats", "az_atoms_out", "az_dump", "az_engine_info_t"]
    probes = {"az_system_desc": ["n_atoms", "xs", "types", "la", "species", "elec_kind", "photons", "seed", "timestep", "efield", "global_ids", "n_atoms_global", "ewald_k", "n_bonds", "bond_types", "bond_type", "n_angles", "angle_types", "angle_type"],
              "az_cuda_opts": ["rebuild_every", "skin", "rank", "pgrid", "nccl_id", "max_neigh"],
              "az_stats": ["engTot", "posMom", "pressure", "specAcBoxPos", "engKin_d", "n_pairs", "engCoul2_d", "engBond", "engAngle_d"],
              "az_engine_info_t": ["cnum", "csize", "pair_blocks", "device_bytes", "list_rows", "capacity"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "aztot_b200.h"', 'int main(void) {']
    for s in structs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f in probes.get(s, []):
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    lines += ['return 0; }']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = dict(zip(out[0::2], map(int, out[1::2])))
    for s in structs:
        cls = getattr(abi, s)
        assert C.sizeof(cls) == got[s], (s, C.sizeof(cls), got[s])
        for f in probes.get(s, []):
            assert getattr(cls, f).offset == got["%s.%s" % (s, f)], (s, f)


def test_no_device_means_an_error_not_a_fallback():
    lib = abi.load_library()
    if lib.az_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import azt
    from aztotmd_b200.engine import AzError, Engine
    with pytest.raises(AzError) as ei:
        Engine(azt.lj_system(ncell=(3, 3, 3)), lib=lib)
    assert ei.value.status == abi.AZ_ERR_NO_DEVICE and "no CPU fallback" in str(ei.value)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "aztotmd_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "az_oracle" not in text and "libaz_oracle" not in text and "oracle/_ref" not in text, os.path.join(dirpath, f)


def _adapter_block():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```cpp\n(// az_adapter\.cpp.*?)```", doc, re.S)
    assert m, "INTEGRATION.md section B lost its adapter block"
    return m.group(1)


def test_reference_adapter_in_the_docs_is_the_checked_file():
    """INTEGRATION.md section B and oracle/az_adapter.cpp hold the same text (the file is what gets compiled below)."""
    src = open(os.path.join(ROOT, "oracle", "az_adapter.cpp")).read()
    marker = "// ---- INTEGRATION.md section B ----\n"
    assert marker in src
    assert src.split(marker, 1)[1] == _adapter_block()


REF_SRC = os.environ.get("AZ_REFERENCE_SRC", "/root/reference/src")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_SRC, "dataStruct.h")), reason="the reference tree is not on this box")
def test_reference_adapter_compiles_against_the_reference_headers():
    """The binding a maintainer would add against the reference's own structs
    (src/dataStruct.h Atoms / Field / Sim / Box / Elec / Spec / VdW, src/temperature.h TStat, src/elec.h tpElec*)
    and include/aztot_b200.h: every field it names exists with a convertible type."""
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", REF_SRC,
           "-I", os.path.join(ROOT, "include"), "-include", "cstdio", os.path.join(ROOT, "oracle", "az_adapter.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
