#!/usr/bin/env bash
# Build the reference's own CUDA program (azTotMD 2.0, src/main.cu) for sm_100a into oracle/_ref/aztotmd_ref.
# TEST / MEASUREMENT INFRASTRUCTURE: used only by profiles/time_reference_cuda.py to time the reference's CUDA
# build on the GPU box beside the new engine.  The sources are compiled from patched TEMPORARY copies under
# /tmp (nothing of the reference enters the repository); the output is git-ignored and travels with gpurun.
#
# Shims (none touches hot-path arithmetic):
#   conio.h / fscanf_s          oracle/shim/conio.h
#   src/cuUtils.cu:86           `return unsigned int (...)` -> C cast (dead function rand1)
#   src/cuElec.cu:11-12,39-90,136-149 and src/main.cu:48: legacy texture<> objects (API removed in CUDA 12):
#                               declarations dropped, init/free_realEwald_tex and real_ewald_tex stubbed (dead code)
#   src/vdw.cpp:308             missing `return 1;` at the end of read_vdw
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AZ_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "build_ref_cuda: $REF not present (GPU box?) - keeping prebuilt files"; exit 0; fi
NVCC="${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}"
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/azrefcuda.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.cu "$REF"/*.cpp "$REF"/*.h "$TMP"/
cp "$HERE/shim/conio.h" "$TMP"/
rm -f "$TMP/main.cpp"   # the serial program's main
python3 - "$TMP" <<'PY'
import os, re, sys
os.chdir(sys.argv[1])
def rw(fn, f):
    s = open(fn, encoding="latin-1").read()
    open(fn, "w", encoding="latin-1").write(f(s))
rw("cuUtils.cu", lambda s: s.replace("return unsigned int ((m_z << 16) + m_w) % 1000;", "return (unsigned int)((m_z << 16) + m_w) % 1000;"))
def vdw(s):
    key = "field->vdws[at2][at1] = &field->pairpots[id];"
    i = s.index(key) + len(key); j = s.index("}", i)
    return s[:j] + "   return 1;\n" + s[j:]
rw("vdw.cpp", vdw)
def elec(s):
    s = re.sub(r"texture<float, 1, cudaReadModeElementType> texCoul\w+;[^\n]*\n", "", s)
    a = s.index("void init_realEwald_tex"); b = s.index("void free_realEwald_tex"); c = s.index("}", b) + 1
    s = s[:a] + "void init_realEwald_tex(cudaMD *md, float mxRange, float alpha) {}\nvoid free_realEwald_tex(cudaMD* md) {}\n" + s[c:]
    a = s.index("__device__ float real_ewald_tex"); b = s.index("__global__ void recip_ewald")
    return s[:a] + "__device__ float real_ewald_tex(float r2, float& r, float chprd, float alpha, float& eng) { return 0.f; }\n\n" + s[b:]
rw("cuElec.cu", elec)
rw("main.cu", lambda s: s.replace("texture<float, 2, cudaReadModeElementType> qProd;", "// texture<float, 2, cudaReadModeElementType> qProd;"))
PY
cd "$TMP"
pids=()
for f in *.cu; do "$NVCC" -O2 -lineinfo -arch=sm_100a -rdc=true -w -Xcompiler -fpermissive -I. -c "$f" -o "${f%.cu}.o" & pids+=($!); done
for f in angles bonds box cell_list ejump elec integrators md_utils out_md rdf sys_init temperature utils vdw; do
  g++ -O2 -std=c++17 -fpermissive -w -I. -c "$f.cpp" -o "$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -arch=sm_100a -rdc=true -o "$OUT/aztotmd_ref" *.o
echo "built $OUT/aztotmd_ref"
