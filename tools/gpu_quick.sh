#!/bin/bash
# parity suite + the three bench lines, no profiling
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout=180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in 1080p_b64 4k_wide_b16 1080p_stress_b64 1080p_step2_b64; do
[ "$wl" = 1080p_step2_b64 ] && EXTRA="--depth-input lowres" || EXTRA=""
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --workload $wl $EXTRA --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
python - gpurun_out/bench_$wl.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
