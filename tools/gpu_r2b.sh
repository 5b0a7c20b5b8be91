#!/usr/bin/env bash
# Round 2, second session: fused-index marcher — parity tests, then A/B against the v9 loop and the GuidanceNet tile heights.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_tree.py tests/test_gpu_render.py tests/test_gpu_frame.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2b.log
tail -15 gpurun_out/pytest_r2b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/ab_render.py fused= v9=RTO_FUSED_INDEX=0 th6=RTO_NET_TILE_H=6 th8=RTO_NET_TILE_H=8 2>&1 | tail -12
