#!/usr/bin/env bash
# Round 2 final single-GPU record: ncu evidence + full bench of the shipped configuration, then compute-sanitizer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 bash tools/gpu_profile_r02.sh > gpurun_out/profile.log 2>&1
grep -E "==|duration" gpurun_out/r02_render_bench_ncu_full.txt | head -4
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','value_reference_protocol','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', round(d['roofline']['frac'],3), d['reference_cuda'].get('speedup_same_protocol'))
print({k:(v.get('value'), v.get('value_reference_protocol'), v.get('e2e')) for k,v in d['configs'].items() if isinstance(v, dict)})
PY
SMOKE_TOOLS=memcheck timeout 900 bash tools/gpu_sanitize.sh > /dev/null 2>&1
cat gpurun_out/sanitizer.txt | grep -E "====|SUMMARY|OK|rror" | head -20
