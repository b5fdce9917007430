#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1x}
timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout=180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 4k_wide_b16 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 1080p_stress_b64 --no-cpu-baseline > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err
tail -30 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
for f in gpurun_out/bench_1080p.json gpurun_out/bench_4k.json gpurun_out/bench_stress.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
if grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then bash tools/gpu_prof.sh $TAG; fi
