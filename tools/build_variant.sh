#!/usr/bin/env bash
# Build a tuning variant of the library with extra nvcc flags: tools/build_variant.sh NAME "-DFOO=1 ..."
# -> build/var_NAME/librtoctree_b200.so (select at run time with RTO_LIB=...).
set -euo pipefail
cd "$(dirname "$0")/.."
name="$1"; flags="${2:-}"
out="build/var_$name"; mkdir -p "$out/obj"
for f in rto_api rto_tree rto_render rto_denoise rto_denoise_tc; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 -Xptxas -v $flags \
       -Iinclude -c rt_octree_b200/csrc/$f.cu -o "$out/obj/$f.o" 2> "$out/obj/$f.ptxas.log" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/librtoctree_b200.so" "$out"/obj/*.o -lcuda
echo "built $out/librtoctree_b200.so"
