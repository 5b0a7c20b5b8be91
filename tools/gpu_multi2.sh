#!/usr/bin/env bash
# usage (under gpurun --gpus N): tools/gpu_multi2.sh N [steps]  -> multi-GPU tests, tile-split check, frame-sharded bench at N, CLI sharding
cd "$(dirname "$0")/.."
N=${1:-2}; K=${2:-200}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
timeout 600 python -m pytest tests/test_cli.py tests/test_sharding.py -m gpu -q --tb=short -p no:cacheprovider -k "sharding or tile_split or pipeline or peer" > gpurun_out/pytest_multi_$N.log 2>&1
tail -3 gpurun_out/pytest_multi_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/tile_split_check.py --width 3840 --height 2160 --depth 9 --frames 12 > gpurun_out/tile_split_$N.json 2> gpurun_out/tile_split_$N.err
echo "tile split exit $?"; tail -1 gpurun_out/tile_split_$N.json | cut -c1-1500; tail -3 gpurun_out/tile_split_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $K --warmup 10 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
    print('N=$N fps %.0f ms/step %.4f protocol %.0f e2e %.0f e2e_f32 %.0f' % (d['value'], d['ms_per_step'], d['value_reference_protocol'], d['e2e']['value'], d['e2e_f32']['value']))
    print(json.dumps(d['configs'])[:1500])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/scale_$N.err').read()[-1500:])
PY
timeout 900 python tools/cli_bench.py --gpus $N --pipe 8 --frames 600 > gpurun_out/cli_bench_$N.json 2> gpurun_out/cli_bench_$N.err; echo "cli bench exit $?"; cat gpurun_out/cli_bench_$N.json | cut -c1-1200
# the product binary's own tile split: 3840x2160, bands over the N GPUs, balanced and static, 60 poses
C=/tmp/rto_cache/cli
for mode in balanced static; do
  if [ "$mode" = static ]; then export RTO_TILE_SPLIT_STATIC=1; else unset RTO_TILE_SPLIT_STATIC; fi
  timeout 600 rt_octree_b200/bin/volrend_headless $C/tree.npz $C/transforms_test.json --options $C/opt.json --ts_module $C/ts_latest.ts \
      -w 3840 -h 2160 --warmup 20 --max_imgs 60 --tile_split --num_gpus $N > gpurun_out/cli_tile_split_${N}_$mode.txt 2>&1
  echo "cli tile split ($mode) exit $?"; grep -E "tile split:|latency:|slowest band" gpurun_out/cli_tile_split_${N}_$mode.txt
done
unset RTO_TILE_SPLIT_STATIC
