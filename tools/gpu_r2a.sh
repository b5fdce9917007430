#!/bin/bash
# round-2 check: GPU tests (incl. the reference-on-CUDA tier), smoke, the default bench line, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout=300 --timeout-method=thread -rs > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?" >> gpurun_out/bench_default.err
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?" >> gpurun_out/bench_ref.err
tail -15 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench_default.err; tail -3 gpurun_out/bench_ref.err
