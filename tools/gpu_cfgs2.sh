for wl in 1080p_step2_b64 1080p_b16_cfg1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-4k --no-f32 --video-frames 0 --e2e-steps 4 > gpurun_out/cfg_$wl.json 2> gpurun_out/cfg_$wl.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload 1080p_b64 --depth-input lowres --no-cpu-baseline --no-4k --no-f32 --video-frames 0 --e2e-steps 4 > gpurun_out/cfg_lowres.json 2> gpurun_out/cfg_lowres.err
python - <<'PY'
import json
for f in ("cfg_1080p_step2_b64","cfg_1080p_b16_cfg1","cfg_lowres"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "fps", round(d["value"]), "ms", round(d["ms_per_step"],4), "kfrac", round(d["roofline"]["frac"],3), {k: round(v,4) for k,v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), d["config"]["workload"][-60:])
    except Exception as e: print(f, "ERR", e)
PY
