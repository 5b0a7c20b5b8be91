#!/usr/bin/env bash
# Writes profiles/<round>_sass_tc.txt: SASS evidence that the GuidanceNet kernel runs on tcgen05 / TMEM / bulk-copy (TMA) units,
# and the marching loop of the production render kernel.  CPU only: cuobjdump reads the built library.
#   tools/sass_evidence.sh r02
set -euo pipefail
cd "$(dirname "$0")/.."
R="${1:-r02}"
LIB=rt_octree_b200/librtoctree_b200.so
OUT=profiles/${R}_sass_tc.txt
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
{
  echo "# cuobjdump -sass $LIB   ($(date -u +%Y-%m-%d), $(nvcc --version | tail -2 | head -1))"
  echo "# sm_100a mnemonics: UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,"
  echo "# UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP.S.G = cp.async.bulk global -> shared (TMA engine, 1-D: weights),"
  echo "# UTMALDG = cp.async.bulk.tensor (tensor-map TMA: the fp32 input tile), SYNCS.* = mbarrier ops"
  echo
  echo "## per kernel: tensor-core / TMEM / TMA / mbarrier instruction counts"
  awk '/Function : /{name=$3} /UTCHMMA|LDTM|UBLKCP|UTCBAR|UTCATOMSWS|UTMALDG|SYNCS\./{ match($0, /(UTCHMMA|LDTM|UBLKCP|UTCBAR|UTCATOMSWS|UTMALDG|SYNCS)[.A-Za-z0-9_]*/); k=name " " substr($0, RSTART, RLENGTH); c[k]++ } END{for (k in c) print c[k], k}' "$TMP" | sort -k2,2 -k3,3 | c++filt
  echo
  echo "## guidance_net_tc_kernel<10>: every UTMALDG / UTCHMMA / UTCBAR / LDTM / UBLKCP / UTCATOMSWS line (address, instruction)"
  awk '/Function : .*guidance_net_tc_kernelILi10E/{p=1; next} /Function : /{p=0} p && /UTMALDG|UTCHMMA|LDTM|UBLKCP|UTCBAR|UTCATOMSWS/ && !/^\s*\/\* 0x/{print}' "$TMP" | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//' | head -80
  echo
  echo "## render_kernel<6,false,16> (production: fused-index marcher, grid level K = 6 = the bench tree): the marching loop,"
  echo "## from the FFMA.SAT position update to the loop branch"
  awk '/Function : _ZN3rto13render_kernelILi6ELb0ELi16EEE/{p=1; next} /Function : /{p=0} p' "$TMP" | grep -v '^\s*/\* 0x' | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//' | awk '/FFMA.SAT/ && !s {s=1} s{print} s && /BSYNC.RECONVERGENT B1/{exit}'
  echo
  echo "## render_kernel<6,false,3> (v9 loop, RTO_FUSED_INDEX=0): the same span, for comparison"
  awk '/Function : _ZN3rto13render_kernelILi6ELb0ELi3EEE/{p=1; next} /Function : /{p=0} p' "$TMP" | grep -v '^\s*/\* 0x' | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//' | awk '/FFMA.SAT/ && !s {s=1} s{print} s && /BRA P1/{exit}'
} > "$OUT"
rm -f "$TMP"
wc -l "$OUT"
