#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-baselines --no-tt > gpurun_out/bench_k20_final.json 2> gpurun_out/bench_k20_final.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k20_final.json'))
print({k:d[k] for k in ('value','value_reference_protocol','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], d['config']==__import__('bench').base_config(1,20))
PY
tail -3 gpurun_out/bench_k20_final.err
