#!/usr/bin/env bash
# GPU A/B of render-kernel build variants (build/var_*/) against the default library; also runs the render parity tests.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
run() { python bench.py --steps 200 --warmup 10 --no-baselines --serial 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: serial fps %.0f render %.4f ms denoise %.4f ms' % (d['value'], d['stage_ms']['render'], d['stage_ms']['denoise']))"; }
run default; run default
shopt -s nullglob
for v in build/var_*/librtoctree_b200.so; do
  n=$(basename $(dirname $v))
  RTO_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "trace or grid" -p no:cacheprovider 2>&1 | tail -1
  RTO_LIB=$PWD/$v run $n; RTO_LIB=$PWD/$v run $n
done
