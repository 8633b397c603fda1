#!/bin/bash
# A/B visit: parity tests of the built library, resident timing of it and of every variant library,
# SM-reservation sweep, launch list of two resident steps, host trace of three resident steps.
# usage (under gpurun): bash tools/gpu_ab.sh <tag> [tests|notests]
tag=${1:-ab}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
  echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
  tail -5 gpurun_out/${tag}_tests.log
fi
bash tools/try_variants.sh 60 5 > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_variants.log
bash tools/knobs.sh > gpurun_out/${tag}_knobs.log 2>&1
cat gpurun_out/${tag}_knobs.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/run_resident.py 60 2 > gpurun_out/${tag}_launches.log 2>&1
ADSB_B200_TRACE=1 timeout 120 python tools/run_resident.py 60 3 > gpurun_out/${tag}_trace.log 2>&1
