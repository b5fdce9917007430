#!/bin/bash
# one full-set capture of selected kernels of one bench step: gpu_prof_k.sh TAG KERNEL_REGEX [bench args]
mkdir -p gpurun_out
TAG=$1; K=$2; shift 2
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-producer --video-frames 0 --no-4k --e2e-steps 1 $*"
timeout -k 10 600 $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 3 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_$TAG.ncu-rep
