#!/usr/bin/env bash
# Round 2, second session: full GPU suite + smoke on the final tree, then the resident-blocks sweep of the render kernel with the v10 loop.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/ab_render.py bps10= bps8=RTO_RENDER_BLOCKS_PER_SM=8 bps6=RTO_RENDER_BLOCKS_PER_SM=6 bps5=RTO_RENDER_BLOCKS_PER_SM=5 2>&1 | tail -9
