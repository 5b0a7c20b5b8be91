#!/usr/bin/env bash
# frames in flight sweep (device-timed value and e2e)
cd "$(dirname "$0")/.."
for p in ${1:-3 4 5 6 4}; do python bench.py --steps 300 --warmup 10 --no-baselines --pipe $p 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pipe $p: fps %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"; done
