#!/usr/bin/env bash
# compute-sanitizer memcheck + racecheck over smoke() (render trace + grid kernels, tcgen05 GuidanceNet, separable filter,
# training forward/backward) and memcheck over tools/sanitize_tree.py (device tree loader incl. the quantised path, both
# brick planes).  Output -> gpurun_out/sanitizer.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  for tool in ${SMOKE_TOOLS:-memcheck racecheck}; do
    echo "==== compute-sanitizer --tool $tool"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -15
  done
  for tool in memcheck racecheck; do
    echo "==== compute-sanitizer --tool $tool tools/sanitize_tree.py"
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_tree.py 2>&1 | grep -v "^$" | tail -8
  done
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
