#!/usr/bin/env bash
# Round 2, second session: N-GPU record of the final state — frame-sharded bench (incl. config 5 and the CLI) + CLI tile split, both read-back modes.
cd "$(dirname "$0")/.."
N=${1:-8}; K=${2:-20}
bash tools/gpu_multi_bench.sh $N $K notests
C=/tmp/rto_cache/cli
for mode in peer band; do
  extra=""; [ "$mode" = band ] && extra="--band_readback"
  timeout 300 rt_octree_b200/bin/volrend_headless $C/tree.npz $C/transforms_test.json --options $C/opt.json --ts_module $C/ts_latest.ts \
      -w 3840 -h 2160 --warmup 20 --max_imgs 60 --tile_split --num_gpus $N $extra > gpurun_out/cli_tile_split_${N}_$mode.txt 2>&1
  echo "cli tile split ($mode) exit $?"; grep -E "tile split:|latency:|slowest band" gpurun_out/cli_tile_split_${N}_$mode.txt
done
