#!/usr/bin/env bash
# Build the TEST-ONLY emulator library: the engine + kernel sources compiled by g++ against
# tests/emu/cuda_emu.h.  Output: tests/emu/_build/libaztot_emu.so (git-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SRC="$ROOT/aztotmd_b200/csrc"
OUT="$HERE/_build"
mkdir -p "$OUT"
CXX="${CXX:-g++}"
FLAGS="-O2 -g -std=c++17 -fPIC -DAZ_EMU -ffp-contract=off -I$HERE -I$SRC -I$ROOT/include -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable"
pids=()
for f in az_engine az_pair_inst_lj az_pair_inst_buck az_pair_inst_surk az_pair_inst_mixed ${AZ_EMU_EXTRA:-}; do
  if [ ! -f "$OUT/$f.o" ] || [ -n "$(find "$SRC" "$HERE" "$ROOT/include" -newer "$OUT/$f.o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) -print -quit)" ]; then
    $CXX $FLAGS -x c++ -c "$SRC/$f.cu" -o "$OUT/$f.o" &
    pids+=($!)
  fi
done
if [ ! -f "$OUT/az_comm_emu.o" ] || [ "$HERE/az_comm_emu.cpp" -nt "$OUT/az_comm_emu.o" ] || [ "$SRC/az_comm.h" -nt "$OUT/az_comm_emu.o" ] || [ "$HERE/cuda_emu.h" -nt "$OUT/az_comm_emu.o" ]; then
  $CXX $FLAGS -I"$SRC" -c "$HERE/az_comm_emu.cpp" -o "$OUT/az_comm_emu.o" &
  pids+=($!)
fi
if [ ! -f "$OUT/cuda_emu.o" ] || [ "$HERE/cuda_emu.cpp" -nt "$OUT/cuda_emu.o" ] || [ "$HERE/cuda_emu.h" -nt "$OUT/cuda_emu.o" ]; then
  $CXX $FLAGS -c "$HERE/cuda_emu.cpp" -o "$OUT/cuda_emu.o" &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX -shared -o "$OUT/libaztot_emu.so" "$OUT"/*.o -lpthread -lm
# the host driver (az_main.cpp + the deck parser) linked against the emulator engine: the process surface of the
# reference (decks in, stat.dat / msd.dat / rdf.dat / revcon.xyz out) can then be tested without a GPU
if [ ! -f "$OUT/aztotmd_emu" ] || [ "$SRC/az_main.cpp" -nt "$OUT/aztotmd_emu" ] || [ "$SRC/az_deck.cpp" -nt "$OUT/aztotmd_emu" ] || [ "$OUT/libaztot_emu.so" -nt "$OUT/aztotmd_emu" ]; then
  $CXX -O2 -std=c++17 -I"$ROOT/include" "$SRC/az_main.cpp" "$SRC/az_deck.cpp" -o "$OUT/aztotmd_emu" -L"$OUT" -laztot_emu -Wl,-rpath,'$ORIGIN' -lpthread -lm
fi
echo "built $OUT/libaztot_emu.so"
