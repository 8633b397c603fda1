#!/bin/bash
# One GPU-box visit: parity tests, resident bench of the built library, pipe microbenchmarks, ncu of k_fused.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
  echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
  tail -5 gpurun_out/${tag}_tests.log
fi
timeout 300 python tools/run_resident.py 60 6 > gpurun_out/${tag}_resident.log 2>&1
tail -4 gpurun_out/${tag}_resident.log
if [ -x tools/ubench/pipes ]; then timeout 120 tools/ubench/pipes > gpurun_out/${tag}_pipes.log 2>&1; cat gpurun_out/${tag}_pipes.log; fi
if [ "${3:-ncu}" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fused|k_gate|k_slice" -s 6 -c 3 -f -o gpurun_out/${tag}_fused \
      python tools/run_resident.py 20 3 > gpurun_out/${tag}_ncu.log 2>&1
  tail -3 gpurun_out/${tag}_ncu.log
fi
