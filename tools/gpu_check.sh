#!/bin/bash
# One gpurun call: GPU parity tests, smoke, microbenchmark, first bench lines.  Everything is wrapped
# in `timeout` so a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 120 ./profiles/microbench/smem_scatter > gpurun_out/microbench.log 2>&1
for mode in 1 2; do
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --scatter-mode $mode --no-cpu-baseline > gpurun_out/bench_1080p_mode$mode.json 2> gpurun_out/bench_1080p_mode$mode.err
done
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --workload 4k_wide_b16 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --workload 1080p_stress_b64 --no-cpu-baseline > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/microbench.log; cat gpurun_out/bench_1080p_mode1.json | cut -c1-600
