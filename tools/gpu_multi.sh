#!/usr/bin/env bash
# usage (under gpurun --gpus N): tools/gpu_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
for n in 1 $N; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 200 --warmup 10 --no-baselines > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python -c "
import json; d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1]); print('N=$n fps %.0f ms/step %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/scale_$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/tile_split_check.py --width 3840 --height 2160 > gpurun_out/tile_split_$N.json 2> gpurun_out/tile_split_$N.err; echo "tile split exit $?"; tail -1 gpurun_out/tile_split_$N.json; tail -3 gpurun_out/tile_split_$N.err
python tools/tile_split_check.py --width 3840 --height 2160 > gpurun_out/tile_split_1.json 2>> gpurun_out/tile_split_$N.err; tail -1 gpurun_out/tile_split_1.json
rt_octree_b200/bin/volrend_headless --help > /dev/null && echo cli-ok
