timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "depth_tail or lowres or producer or fp32_depth_through" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --workload 1080p_b64 --depth-input lowres --no-cpu-baseline --no-4k --no-f32 --video-frames 0 --e2e-steps 2 > gpurun_out/cfg_lowres.json 2> gpurun_out/cfg_lowres.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/cfg_lowres.json")); print("lowres fps", round(d["value"]), "ms", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
PY
