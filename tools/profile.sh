#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list of the bench command + one full capture of the top kernel.
# Usage (under gpurun): bash tools/profile.sh <tag> [bodies]
TAG=${1:-r01}; BODIES=${2:-200000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --bodies $BODIES --cpu-sample 2000 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chat_ -s 3 -c 3 -f -o gpurun_out/chat_${TAG} \
    python bench.py --steps 1 --warmup 1 --bodies $BODIES --cpu-sample 2000 --skip-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
