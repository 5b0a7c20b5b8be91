#!/usr/bin/env bash
# GPU tuning helper: frame time vs resident blocks/SM of the persistent render kernel.
cd "$(dirname "$0")/.."
for b in ${*:-3 4 5 6 8 10}; do
  RTO_RENDER_BLOCKS_PER_SM=$b python bench.py --steps 100 --warmup 5 --no-baselines 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/SM $b: fps %.0f render %.3f ms denoise %.3f ms e2e %.0f' % (d['value'], d['stage_ms']['render'], d['stage_ms']['denoise'], d['e2e']['value']))"
done
