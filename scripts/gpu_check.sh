#!/bin/bash
# One GPU-box round trip: parity tests, smoke, a short bench, and the ncu launch list.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
  echo "pytest exit $?" >> gpurun_out/tests.log
  tail -n 60 gpurun_out/tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -n 5 gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --blobs 512 --no-cpu-baseline --no-passes \
    > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_bench.log
fi
