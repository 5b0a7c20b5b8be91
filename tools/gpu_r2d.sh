#!/usr/bin/env bash
# Round 2, second session, call 3: denoiser / CLI / frame / sharding tests after the filter anchoring fix, ncu evidence, full bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_denoise.py tests/test_gpu_frame.py tests/test_cli.py tests/test_sharding.py tests/test_gpu_vs_reference.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_r2d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2d.log
tail -6 gpurun_out/pytest_r2d.log
timeout 1500 bash tools/gpu_profile_r02.sh > gpurun_out/profile.log 2>&1
tail -12 gpurun_out/profile.log
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','value_reference_protocol','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('value_per_frame_calls'), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), d.get('reference_cuda'))
PY
tail -3 gpurun_out/bench.err
