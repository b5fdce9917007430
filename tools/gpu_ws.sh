#!/bin/bash
# warp-kernel iteration: the parity tests that exercise it, then the three device-resident benches
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard.py -m gpu -x -q --tb=short -p no:cacheprovider --timeout=300 --timeout-method=thread -k "${TESTS:-small_cases or wide or split or guard or full_batch or edge}" > gpurun_out/pytest_ws.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ws.log
tail -4 gpurun_out/pytest_ws.log
for wl in 1080p_b64 4k_wide_b16 ${EXTRA_WL}; do
  timeout -k 10 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-4k --video-frames 0 --e2e-steps 2 > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err
  python - "$wl" <<'PY'
import json,sys
f="gpurun_out/q_%s.json"%sys.argv[1]
try:
    d=json.load(open(f)); print(sys.argv[1], "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "kfrac", round(d["roofline"]["frac"],3), "stage_frac", round(d["roofline"]["stage_frac"],3), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), "same", d["e2e_frames_equal"])
except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
PY
done
for o in $AB_OPTS; do
  timeout -k 10 600 python bench.py --steps 10 --warmup 3 --workload ${AB_WL:-1080p_b64} --no-cpu-baseline --no-4k --video-frames 0 --e2e-steps 2 --opt $o > gpurun_out/q_ab_$o.json 2> gpurun_out/q_ab_$o.err
  python - "$o" <<'PY'
import json,sys
f="gpurun_out/q_ab_%s.json"%sys.argv[1]
try:
    d=json.load(open(f)); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
done
