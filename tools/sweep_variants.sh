#!/usr/bin/env bash
# GPU tuning helper: A/B the render kernel variants (env-selected) on the bench workload (pipelined and serial).
cd "$(dirname "$0")/.."
run() { python bench.py --steps 200 --warmup 10 --no-baselines $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: fps %.0f render %.3f ms denoise %.3f ms' % (d['value'], d['stage_ms']['render'], d['stage_ms']['denoise']))"; }
run "A default (4 warps/block, 8 blocks)"
RTO_LIB=$PWD/build/varB/librtoctree_b200.so run "B 8 warps/block (32x8 super-tiles), 4 blocks"
RTO_LIB=$PWD/build/varB/librtoctree_b200.so RTO_RENDER_BLOCKS_PER_SM=3 run "B 8 warps/block, 3 blocks"
