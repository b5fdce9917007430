#!/bin/bash
# 1/2/4/8-GPU line on one 8-GPU box (gpurun --gpus 8): the default bench (device-resident value, e2e through submit/collect,
# aggregate PCIe probe, sharded video leg), short; one JSON line per N in gpurun_out/scale8_N.json
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/scale8_topo.txt 2>&1; nproc >> gpurun_out/scale8_topo.txt; numactl -H >> gpurun_out/scale8_topo.txt 2>&1; free -g >> gpurun_out/scale8_topo.txt
for N in ${NS:-1 2 4 8}; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-4k --e2e-steps 5 "$@" > gpurun_out/scale8_$N.json 2> gpurun_out/scale8_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-4k --e2e-steps 5 "$@" > gpurun_out/scale8_$N.json 2> gpurun_out/scale8_$N.err
  fi
  tail -1 gpurun_out/scale8_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print(d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(e['value']), 'pcie', {k: round(v,1) for k,v in e['pcie'].items() if k.endswith('gbs')}, 'ceiling', round(e['pcie']['dma_ceiling_fps']), 'frac', round(e['frac_of_dma_ceiling'],3), 'lowres', round(e['lowres_depth_value']), 'devdepth', round(e['device_depth_value']), 'video', round(d['video']['frames_per_sec']) if d.get('video') else None)" || tail -5 gpurun_out/scale8_$N.err
done
