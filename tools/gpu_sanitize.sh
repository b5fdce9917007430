#!/bin/bash
# compute-sanitizer over tools/sanitize_case.py: memcheck, racecheck, synccheck, initcheck; logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
TAG=${1:-r02}
timeout -k 10 300 python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -20 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck synccheck racecheck initcheck; do
  extra=""
  [ $tool = racecheck ] && export SAN_CASES=small_a
  timeout -k 10 1500 compute-sanitizer --tool $tool $extra --print-limit 30 python tools/sanitize_case.py > gpurun_out/sanitize_${TAG}_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_${TAG}_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit " gpurun_out/sanitize_${TAG}_$tool.log | tail -3
done
