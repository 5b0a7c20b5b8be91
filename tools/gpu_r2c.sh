#!/usr/bin/env bash
# Round 2, second session, call 2: full GPU suite on the fused-index marcher + trimmed denoiser, A/B numbers, short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/ab_render.py new= 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-baselines --no-tt > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k20.json'))
print({k:d[k] for k in ('value','value_reference_protocol','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('value_per_frame_calls'), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3))
print({k:(v.get('value'), v.get('e2e')) for k,v in d['configs'].items() if isinstance(v, dict)})
PY
tail -3 gpurun_out/bench_k20.err
