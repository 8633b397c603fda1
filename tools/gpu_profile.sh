#!/bin/bash
# Profile evidence of one build: bench line, ncu launch list of the same command, ncu --set full of the three
# hot-path kernels at bench size.  usage (under gpurun): bash tools/gpu_profile.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fused|k_gate|k_slice" -s 9 -c 3 -f \
    -o gpurun_out/${tag}_kernels python tools/run_resident.py 60 4 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
python tools/showbench.py gpurun_out/${tag}_bench.json | cut -c1-400
