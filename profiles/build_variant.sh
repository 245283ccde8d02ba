#!/bin/bash
# usage: profiles/build_variant.sh NAME [-DMACRO=..]...   -> aztotmd_b200/lib/variants/NAME/libaztot_b200.so
# A/B builds of the engine for one gpurun job (the job copies a variant over lib/libaztot_b200.so on the box).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=aztotmd_b200/lib/variants/$name
mkdir -p $out/obj
C=${AZ_VARIANT_SRC:-aztotmd_b200/csrc}
pids=()
for f in az_engine az_pair_inst_lj az_pair_inst_buck az_pair_inst_surk az_pair_inst_mixed; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include -I $C "$@" -c $C/$f.cu -o $out/obj/$f.o &
  pids+=($!)
done
for f in az_deck az_comm_nccl; do
  g++ -O2 -std=c++17 -fPIC -I include -I $C -I /usr/local/cuda/include -c $C/$f.cpp -o $out/obj/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libaztot_b200.so $out/obj/*.o -ldl -Xlinker -soname,libaztot_b200.so
echo built $out/libaztot_b200.so
