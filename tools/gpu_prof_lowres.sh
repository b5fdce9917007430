CMD="python bench.py --workload 1080p_b64 --depth-input lowres --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-4k --no-f32 --video-frames 0"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_depth_lowres_tiled -s 67 -c 1 -o gpurun_out/prof_r02d_lowres $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
