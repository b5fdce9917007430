#!/bin/bash
# fused-route bring-up: debug diff first (cheap), then the parity suite, smoke and bench lines
mkdir -p gpurun_out
timeout -k 10 300 python tools/debug_fused.py > gpurun_out/debug_fused.log 2>&1
echo "debug exit $?" >> gpurun_out/debug_fused.log
timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout=180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 4k_wide_b16 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 1080p_stress_b64 --no-cpu-baseline > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err
cat gpurun_out/debug_fused.log | tail -40; tail -30 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
for f in gpurun_out/bench_1080p.json gpurun_out/bench_4k.json gpurun_out/bench_stress.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "frac", d["roofline"]["frac"], d["stage_ms_per_step"], "e2e", round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
