#!/usr/bin/env bash
# N-GPU check (tools/gpu_multi_bench.sh N K [notests]): multi-GPU tests + the frame-sharded bench (rto_frame_sequence e2e, 4K tile split, CLI --num_gpus) at N = 2.
cd "$(dirname "$0")/.."
N=${1:-2}; K=${2:-20}
mkdir -p gpurun_out
[ "${3:-tests}" = notests ] || timeout 600 python -m pytest tests/test_cli.py tests/test_sharding.py -m gpu -q --tb=short -p no:cacheprovider -k "sharding or tile_split or peer" > gpurun_out/pytest_multi_$N.log 2>&1
[ "${3:-tests}" = notests ] || tail -3 gpurun_out/pytest_multi_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $K --warmup 3 --no-baselines > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
    print('N=$N fps %.0f ms/step %.4f protocol %.0f e2e %.0f (per-frame calls %.0f) e2e_f32 %.0f' % (d['value'], d['ms_per_step'], d['value_reference_protocol'], d['e2e']['value'], d['e2e']['value_per_frame_calls'], d['e2e_f32']['value']))
    print(json.dumps(d['configs'])[:1800])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/scale_$N.err').read()[-1500:])
PY
