#!/usr/bin/env bash
# Build oracle/_ref/libref_device.so: the reference's own __device__ functions of the hot path compiled as HOST code
# (see oracle/ref_device_harness.cpp).  TEST INFRASTRUCTURE.  The function bodies are cut verbatim out of the
# reference's .cu files into a temporary include file under /tmp; nothing of the reference enters the repository.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AZ_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "build_ref_device: $REF not present (GPU box?) - keeping prebuilt files"; exit 0; fi
CUDA_INC="${CUDA_INC:-/usr/local/cuda/include}"
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/azrefdev.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
python3 - "$REF" "$TMP/ref_device_funcs.inc" <<'PY'
import re, sys
ref, out = sys.argv[1], sys.argv[2]
want = {
    "cuUtils.cu": ["float3_sqr", "float3_length"],
    "cuVdW.cu": ["cu_fe_lj", "cu_fer_lj", "cu_fe_buck", "cu_fer_buck", "cu_fe_bmh", "cu_fer_bmh", "cu_fe_elin", "cu_fer_elin",
                 "cu_fe_einv", "cu_fer_einv", "surk_pot"],
    "cuElec.cu": ["direct_coul", "real_ewald", "fennel"],
    "cuTemp.cu": ["rand_uvect", "get_angled_vector", "adsorb_rand_photon", "radiate_photon3", "before_nose", "after_nose"],
    "cuMDfunc.cu": ["put_periodic", "delta_periodic"],
    "cuBonds.cu": ["bond_harm", "bond_morse", "bond_pedone", "bond_buck", "bond_e6812"],   # SURVEY 8f-4 groundwork
    "cuAngles.cu": ["angle_hcos"],
    "cuSort.cu": ["count_cell"],
    "cuCellList.cu": ["split_cells", "cell_dist_int"],   # host functions of the CUDA build
}
def strip_comments(s):  # only to find the braces; the emitted text is the original
    s = re.sub(r"/\*.*?\*/", lambda m: " " * len(m.group(0)), s, flags=re.S)
    return re.sub(r"//[^\n]*", lambda m: " " * len(m.group(0)), s)
chunks = []
for fn, names in want.items():
    src = open("%s/%s" % (ref, fn), encoding="latin-1").read()
    bare = strip_comments(src)
    if fn == "cuTemp.cu":  # the four physical constants in front of rand_uvect
        for m in re.finditer(r"^__constant__ const float (revLight|Light|revPlank|numPi)[^\n]*\n", src, flags=re.M):
            chunks.append(m.group(0))
    for name in names:
        m = None
        for cand in re.finditer(r"^(?:__device__\s+|__global__\s+)?[\w\*]+[\w\s\*]*?\b%s\s*\(" % re.escape(name), bare, flags=re.M):
            semi = bare.find(";", cand.end())
            if semi < 0 or bare.index("{", cand.end()) < semi:  # a definition, not a forward declaration
                m = cand
                break
        if not m:
            raise SystemExit("function %s not found in %s" % (name, fn))
        i = bare.index("{", m.end())
        depth, j = 0, i
        while True:
            c = bare[j]
            if c == "{": depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0: break
            j += 1
        chunks.append("// ---- %s : %s\n" % (fn, name) + src[m.start():j + 1] + "\n")
text = "\n".join(chunks)
# one local variable of surk_pot is spelled with a Cyrillic capital (cp1251 byte 0xD1): g++ rejects it in an identifier.
# Every non-ASCII byte becomes 'X' (a consistent rename; the garbled comments do not matter).
text = re.sub(r"[\x80-\xff]", "X", text)
open(out, "w", encoding="latin-1").write(text)
PY
g++ -O2 -ffp-contract=off -fPIC -shared -w -fpermissive -std=c++17 -finput-charset=latin1 -I"$TMP" -I"$REF" -I"$CUDA_INC" \
    "$HERE/ref_device_harness.cpp" -o "$OUT/libref_device.so" -lm
echo "built $OUT/libref_device.so"
