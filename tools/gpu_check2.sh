#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --scatter-mode 1 --no-cpu-baseline > gpurun_out/bench_1080p_mode1.json 2> gpurun_out/bench_1080p_mode1.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 4k_wide_b16 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --workload 1080p_stress_b64 --no-cpu-baseline > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
if grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then
  timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 &&
  timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
  timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain2.log 2>&1 &&
  timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:'k_warp_rows|k_blur_tiles|k_depth_full' -s 9 -c 3 -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
