#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for b in 8 7 6 5; do for p in 4 6; do
  RTO_RENDER_BLOCKS_PER_SM=$b python bench.py --steps 300 --warmup 10 --no-baselines --pipe $p 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks $b pipe $p: fps %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done; done
