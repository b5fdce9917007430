#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "exact fraction|passed|failed" gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain2.log 2>&1 &&
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:'k_warp_rows|k_blur_tiles|k_depth_full' -s 9 -c 3 -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; tail -3 gpurun_out/ncu_launches.log
