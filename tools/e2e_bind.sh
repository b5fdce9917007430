#!/bin/bash
# end-to-end numbers with and without NUMA-local CPU binding (one GPU)
for f in "--bind" ""; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 5 $f 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$f', d['config']['cpu_binding'], round(d['value']), {k: (round(v) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','lowres_depth_value','per_frame_call_fps')})"
done
ls /sys/bus/pci/devices | head -3; nvidia-smi topo -m 2>/dev/null | head -8; nproc
