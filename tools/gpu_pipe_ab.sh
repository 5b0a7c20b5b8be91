#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for p in 2 3 4 2 3; do python bench.py --steps 300 --warmup 10 --no-baselines --pipe $p 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pipe $p: fps %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"; done
