#!/bin/bash
# one bench line per BASELINE.json config that fits one GPU (configs[0..3]); configs[4] is tools/gpu_scale.sh
mkdir -p gpurun_out
for spec in "1080p_b16_cfg1:full" "1080p_b64:full" "1080p_b64:lowres" "4k_wide_b16:full" "1080p_step2_b64:full" "1080p_step2_b64:lowres" "1080p_stress_b64:full"; do
wl=${spec%%:*}; di=${spec##*:}
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --workload $wl --depth-input $di --no-cpu-baseline --e2e-steps 3 > gpurun_out/cfg_${wl}_$di.json 2> gpurun_out/cfg_${wl}_$di.err
python - gpurun_out/cfg_${wl}_$di.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('cfg_')[1], "| fps", round(d["value"]), "| ms/step", round(d["ms_per_step"],3), "| frac", round(d["roofline"]["frac"],3), "|", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "| e2e", round(d["e2e"]["value"]), "|", d["config"]["workload"].split("depth (")[-1])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
