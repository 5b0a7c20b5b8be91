#!/usr/bin/env bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 300 --warmup 10 --no-baselines $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: fps %.0f e2e %.0f render %.4f' % (d['value'], d['e2e']['value'], d['stage_ms']['render']))"; }
run "default 8 blocks"
for v in ${VARS:-mb10:10 mb12:12 mb12:11}; do n=${v%%:*}; b=${v##*:}
  RTO_LIB=$PWD/build/var_$n/librtoctree_b200.so timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "grid" -p no:cacheprovider 2>&1 | tail -1
  RTO_LIB=$PWD/build/var_$n/librtoctree_b200.so RTO_RENDER_BLOCKS_PER_SM=$b run "$n blocks=$b"
done
run "default 8 blocks"
