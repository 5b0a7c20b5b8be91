#!/usr/bin/env bash
# resident blocks per SM of the render kernel (v10 loop), finer sweep + the 64-register build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/ab_render.py bps8=RTO_RENDER_BLOCKS_PER_SM=8 bps7=RTO_RENDER_BLOCKS_PER_SM=7 bps9=RTO_RENDER_BLOCKS_PER_SM=9 mb8=RTO_LIB=$PWD/build/var_mb8/librtoctree_b200.so,RTO_RENDER_BLOCKS_PER_SM=8 2>&1 | tail -9
