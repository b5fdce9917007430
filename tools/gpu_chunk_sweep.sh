#!/bin/bash
# e2e frames/s of the headline workload for several host chunk sizes
for hc in 2 4 8 16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-4k --no-f32 --video-frames 0 --host-chunk $hc > gpurun_out/hc_$hc.json 2> gpurun_out/hc_$hc.err
  python - "$hc" <<'PY'
import json,sys
d=json.load(open("gpurun_out/hc_%s.json"%sys.argv[1])); e=d["e2e"]
print("host_chunk", sys.argv[1], "e2e", round(e["value"]), "ceil", round(e["pcie"]["dma_ceiling_fps"]), "frac", round(e["frac_of_dma_ceiling"],3), "lowres", round(e["lowres_depth_value"]), "dev", round(e["device_depth_value"]))
PY
done
