#!/usr/bin/env bash
# Runs on the GPU box under gpurun: tests, smoke, a short bench, and the ncu passes.  Everything lands in gpurun_out/.
# usage: tools/gpu_check.sh [tests|bench|ncu|all]...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-all}"
has() { [[ " $what " == *" $1 "* || " $what " == *" all "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -40 gpurun_out/pytest_gpu.log
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log
fi
if has bench; then
  timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if has benchq; then
  timeout 600 python bench.py --steps 200 --warmup 10 --no-baselines > gpurun_out/benchq.json 2> gpurun_out/benchq.err; echo "benchq exit $?"
  python - <<'PY'
import json
d=json.load(open('gpurun_out/benchq.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms','value_l2_flushed','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3))
PY
  tail -3 gpurun_out/benchq.err
fi
if has ncur; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o gpurun_out/prof_render \
      python bench.py --steps 12 --warmup 3 --no-baselines > gpurun_out/ncu_render.log 2>&1
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 12 --warmup 3 --no-baselines > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o gpurun_out/prof_render \
      python bench.py --steps 12 --warmup 3 --no-baselines > gpurun_out/ncu_render.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"guidance_net|filter_sep|filter_kernel" -s 6 -c 2 -f -o gpurun_out/prof_denoise \
      python bench.py --steps 12 --warmup 3 --no-baselines > gpurun_out/ncu_denoise.log 2>&1
  ls -la gpurun_out
fi
