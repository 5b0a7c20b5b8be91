#!/usr/bin/env bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 300 --warmup 10 --no-baselines $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: fps %.0f e2e %.0f render %.4f denoise %.4f' % (d['value'], d['e2e']['value'], d['stage_ms']['render'], d['stage_ms']['denoise']))"; }
run default; run default
RTO_L2_PERSIST=1 run "default+persist"
RTO_LIB=$PWD/build/var_hints/librtoctree_b200.so run hints
RTO_LIB=$PWD/build/var_hints/librtoctree_b200.so run hints
RTO_L2_PERSIST=1 RTO_LIB=$PWD/build/var_hints/librtoctree_b200.so run "hints+persist"
M=lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum
for lib in rt_octree_b200/librtoctree_b200.so build/var_hints/librtoctree_b200.so; do
  RTO_LIB=$PWD/$lib timeout 600 ncu --metrics $M --clock-control none -k regex:render_kernel -s 20 -c 2 --csv python bench.py --steps 30 --warmup 5 --no-baselines 2>/dev/null | grep -E "render_kernel" | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -8
done
