#!/usr/bin/env bash
# Build the reference's own serial CPU implementation (aztot_serial's sources) into oracle/_ref/.
# TEST INFRASTRUCTURE: the outputs are git-ignored, travel to the GPU box, and are only used by
# tests/ (oracle pinning) and bench.py (--impl reference / cpu_baseline).
#
# The sources are compiled where they lie under $REF (default /root/reference/src); nothing is
# copied into the repository.  Two files cannot be compiled verbatim on Linux/g++:
#   src/vdw.cpp      read_vdw() falls off the end of an int function (line 308) -> UB at -O2;
#                    a patched copy in a temporary directory gets "return 1;" appended.
#   src/main.cpp     not used: oracle/ref_serial_harness.cpp re-states its loop (and main.cpp:212
#                    calls free_md with the wrong arity).
# <conio.h> and fscanf_s come from oracle/shim/conio.h.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AZ_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "build_ref: $REF not present (GPU box?) - keeping prebuilt files"; exit 0; fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/azref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
CXX="${CXX:-g++}"
FLAGS="-O2 -std=c++17 -fpermissive -w -I$HERE/shim -I$REF"
# patched copy of vdw.cpp: add the missing return at the end of read_vdw
python3 - "$REF/vdw.cpp" "$TMP/vdw.cpp" <<'PY'
import sys, re
src = open(sys.argv[1], encoding="latin-1").read()
key = "field->vdws[at2][at1] = &field->pairpots[id];"
i = src.index(key) + len(key)
j = src.index("}", i)
src = src[:j] + "   return 1;\n" + src[j:]
open(sys.argv[2], "w", encoding="latin-1").write(src)
PY
OBJS=""
for f in angles bonds box cell_list ejump elec integrators md_utils out_md rdf sys_init temperature utils; do
  $CXX $FLAGS -c "$REF/$f.cpp" -o "$TMP/$f.o"
  OBJS="$OBJS $TMP/$f.o"
done
$CXX $FLAGS -c "$TMP/vdw.cpp" -o "$TMP/vdw.o"
$CXX $FLAGS -c "$HERE/ref_serial_harness.cpp" -o "$TMP/harness.o"
$CXX -O2 -o "$OUT/ref_serial_harness" "$TMP/harness.o" $OBJS "$TMP/vdw.o" -lm
echo "built $OUT/ref_serial_harness"
