#!/bin/bash
# Diagnostic A/B builds: the library compiled with extra macros into tools/_alt/libagrl_b200_<name>.so (git-ignored, travels
# with gpurun), selected per process by tools/head_variants.py's HV_LIB.  The product library is not touched.
#   usage: bash tools/build_alt.sh NAME -DMACRO[=v] ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/tools/_alt; obj=$out/obj_$name
mkdir -p "$obj"
for f in "$root"/agrl/pytorch_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -Xfatbin -compress-all -I "$root/include" "$@" -c "$f" -o "$obj/$(basename "${f%.cu}").o" &
done
wait
nvcc -shared -o "$out/libagrl_b200_$name.so" "$obj"/*.o -cudart static -Xcompiler -fPIC
echo "$out/libagrl_b200_$name.so"
