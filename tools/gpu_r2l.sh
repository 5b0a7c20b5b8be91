#!/usr/bin/env bash
# shared-memory carve-out of the render kernel sized for the resident blocks (more L1 for the brick planes) vs the driver's choice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/ab_render.py auto= drv=RTO_SMEM_CARVEOUT=0 c64=RTO_SMEM_CARVEOUT=57 2>&1 | tail -7
