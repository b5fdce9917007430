#!/bin/bash
# weak-scaling run on one multi-GPU box: bash tools/gpu_scale.sh "1 2 4 8" [extra bench.py flags]
# (launch with gpurun --gpus 8); one JSON line per GPU count in gpurun_out/scale_N.json
mkdir -p gpurun_out
NS=${1:-"1 2 4 8"}; shift
for N in $NS; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  tail -1 gpurun_out/scale_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d.get('video'))"
done
