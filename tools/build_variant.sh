#!/usr/bin/env bash
# Build a tuning variant of the library with extra nvcc flags: tools/build_variant.sh NAME "-DFOO=1 ..."
# -> build/var_NAME/librtoctree_b200.so (select at run time with RTO_LIB=...).
set -euo pipefail
cd "$(dirname "$0")/.."
name="$1"; flags="${2:-}"
out="$PWD/build/var_$name"; mkdir -p "$out/obj"
make -C rt_octree_b200/csrc -j8 OBJDIR="$out/obj" LIB="$out/librtoctree_b200.so" EXTRA_NVFLAGS="$flags"
echo "built $out/librtoctree_b200.so"
