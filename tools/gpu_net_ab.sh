#!/usr/bin/env bash
# GPU A/B of the GuidanceNet tile heights (RTO_NET_TILE_H): parity tests + per-kernel ncu times + pipelined bench
cd "$(dirname "$0")/.."
for th in ${1:-6 8 14}; do
  echo "== RTO_NET_TILE_H=$th"
  RTO_NET_TILE_H=$th timeout 600 python -m pytest tests/test_gpu_denoise.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -1
  RTO_NET_TILE_H=$th timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"guidance_net" -s 8 -c 3 --csv python bench.py --steps 12 --warmup 3 --no-baselines 2>/dev/null | grep -E "guidance" | awk -F'"' '{print $(NF-1)}' | tr '\n' ' '; echo
  RTO_NET_TILE_H=$th timeout 300 bash tools/gpu_check.sh benchq 2>&1 | tail -1 | cut -c1-200
done
