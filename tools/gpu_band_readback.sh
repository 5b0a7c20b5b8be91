#!/usr/bin/env bash
# Round 2, second session: band read-back of the tile split (tests + 4K latency, peer-direct assembly vs per-GPU band copies) on N GPUs.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_frame.py -m gpu -q --tb=short -p no:cacheprovider -k "tile_split or image_target or sequence" > gpurun_out/pytest_band_$N.log 2>&1
tail -4 gpurun_out/pytest_band_$N.log
python - <<'PY'
import os, sys
sys.path.insert(0, "tools")
import cli_bench
files = cli_bench.workload_files(os.path.join(cli_bench.bench.CACHE, "cli"))
print(files)
PY
C=/tmp/rto_cache/cli
for mode in peer band; do
  extra=""; [ "$mode" = band ] && extra="--band_readback"
  timeout 600 rt_octree_b200/bin/volrend_headless $C/tree.npz $C/transforms_test.json --options $C/opt.json --ts_module $C/ts_latest.ts \
      -w 3840 -h 2160 --warmup 20 --max_imgs 60 --tile_split --num_gpus $N $extra > gpurun_out/cli_tile_split_${N}_$mode.txt 2>&1
  echo "cli tile split ($mode) exit $?"; grep -E "tile split:|latency:|slowest band" gpurun_out/cli_tile_split_${N}_$mode.txt
done
