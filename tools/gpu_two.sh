#!/bin/bash
# GPU tests + the 1080p_b64 and 4k_wide_b16 bench lines (device-resident numbers only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --timeout=300 --timeout-method=thread 2>&1 | tail -5
for wl in 1080p_b64 4k_wide_b16; do
  python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/two_$wl.log 2>&1
  tail -1 gpurun_out/two_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:12], round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
