#!/usr/bin/env bash
# GPU tuning helper: A/B the render kernel variants (env-selected) on the bench workload.
cd "$(dirname "$0")/.."
run() { python bench.py --steps 100 --warmup 5 --no-baselines 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: fps %.0f render %.3f ms denoise %.3f ms' % (d['value'], d['stage_ms']['render'], d['stage_ms']['denoise']))"; }
run "A default"
RTO_L2_PERSIST=1 run "A + L2 persisting window on bricks"
[ -f build/varB/librtoctree_b200.so ] && RTO_LIB=$PWD/build/varB/librtoctree_b200.so run "B branch-free lookup"
RTO_DISABLE_GRID=1 run "tree walker (grid disabled)"
