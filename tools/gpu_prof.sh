#!/bin/bash
# launch list + one full-set capture of the main kernels of one bench step: gpu_prof.sh TAG [WORKLOAD]
mkdir -p gpurun_out
TAG=${1:-r1x}
WL=${2:-1080p_b64}
EXTRA=${3:-}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-4k --video-frames 0 $EXTRA"
K='k_depth_pass|k_build_tables|k_warp_fused|k_warp_ws|k_blur_holes|k_blur_sep|k_blur_band|k_band_list|k_band_map|k_word_list|k_blur_commit'
timeout -k 10 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 18 -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 18 -c 6 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; grep -c k_ gpurun_out/launches_$TAG.csv
