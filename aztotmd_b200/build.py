"""In-tree nvcc build of the B200 engine (sm_100a only).

``python -m aztotmd_b200.build`` compiles aztotmd_b200/csrc/*.cu into
``aztotmd_b200/lib/libaztot_b200.so`` (C ABI of include/aztot_b200.h) and the host driver
``aztotmd_b200/lib/aztotmd``.  nvcc cross-compiles without a GPU.  The equivalent CMake project
(CMakeLists.txt at the repository root) is the Linux replacement of the reference's MSVC-only
``src/aztotmd.vcxproj``; this script exists so that the build check needs nothing but nvcc.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC]

CU_SOURCES = ["az_engine.cu", "az_pair_inst_lj.cu", "az_pair_inst_buck.cu", "az_pair_inst_surk.cu",
              "az_pair_inst_mixed.cu"]
HOST_SOURCES = ["az_deck.cpp", "az_comm_nccl.cpp"]  # NCCL is loaded with dlopen at the first multi-GPU create


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without CUDA")
    return p


def _newer(src_deps, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_deps)


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "aztot_b200.h"))
    return deps


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((" ".join(cmd), r.stdout + r.stderr))
    if r.returncode != 0:
        raise RuntimeError("build step failed:\n%s\n%s" % (" ".join(cmd), r.stdout + r.stderr))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if anything is out of date; returns the path of libaztot_b200.so."""
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = nvcc_path()
    deps = _deps()
    log = []
    objs = []
    jobs = []
    for src in CU_SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(deps, obj):
            jobs.append([nvcc, *ARCH, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj])
    for src in HOST_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJDIR, src.replace(".cpp", ".o"))
        objs.append(obj)
        if force or _newer(deps, obj):
            jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                         "-I", os.path.join(os.path.dirname(os.path.dirname(nvcc)), "include"), "-c", path, "-o", obj])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(lambda c: _run(c, log), jobs))
    lib = os.path.join(LIBDIR, "libaztot_b200.so")
    if jobs or not os.path.exists(lib):
        _run([nvcc, *ARCH, "-shared", "-o", lib, *objs, "-ldl", "-Xlinker", "-soname,libaztot_b200.so"], log)
    main_src = os.path.join(CSRC, "az_main.cpp")
    exe = os.path.join(LIBDIR, "aztotmd")
    if os.path.exists(main_src) and (jobs or _newer([main_src], exe)):
        _run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), main_src, "-o", exe,
              "-L", LIBDIR, "-laztot_b200", "-Wl,-rpath,$ORIGIN"], log)
    if verbose:
        for cmd, out in log:
            print(cmd)
            print(out)
    with open(os.path.join(LIBDIR, "build.log"), "a" if not force else "w") as f:
        for cmd, out in log:
            f.write(cmd + "\n" + out + "\n")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
