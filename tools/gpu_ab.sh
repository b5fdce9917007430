#!/bin/bash
# A/B runs of the device-resident warp stage: gpu_ab.sh "label:opts" ... (opts = extra bench.py arguments)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout=300 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for spec in "$@"; do
  label=${spec%%:*}; opts=${spec#*:}
  timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-producer --video-frames 0 --no-4k --e2e-steps 2 $opts > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - "$label" <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{l}.json")); print(l, "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "(with events", round(d["ms_per_step_with_stage_events"],4), ") stage_frac", round(d["roofline"]["stage_frac"],3), "kfrac", round(d["roofline"]["frac"],3), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), d["e2e_frames_equal"])
except Exception as e: print(l, "ERR", e, open(f"gpurun_out/ab_{l}.err").read()[-1500:])
PY
done
