#!/usr/bin/env bash
# usage (under gpurun --gpus N): tools/gpu_multi.sh "1 2 4 8"   -> frame-sharded bench at each N + 4K tile split at max N
cd "$(dirname "$0")/.."
LIST=${1:-"1 2"}
mkdir -p gpurun_out
MAXN=1
for n in $LIST; do
  MAXN=$n
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 200 --warmup 10 --no-baselines > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python -c "
import json; d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1]); print('N=$n fps %.0f ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/scale_$n.err
done
if [ "$MAXN" != 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXN --master-addr 127.0.0.1 --master-port 29534 tools/tile_split_check.py --width 3840 --height 2160 > gpurun_out/tile_split_$MAXN.json 2> gpurun_out/tile_split_$MAXN.err; echo "tile split exit $?"; tail -1 gpurun_out/tile_split_$MAXN.json
fi
python tools/tile_split_check.py --width 3840 --height 2160 > gpurun_out/tile_split_1.json 2> gpurun_out/tile_split_1.err; tail -1 gpurun_out/tile_split_1.json
