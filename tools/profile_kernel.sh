#!/bin/bash
# one full ncu capture of kernels matching $1 (regex) from a short resident-only bench; report -> gpurun_out/$2.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c 1 -f -o gpurun_out/$2 \
    python bench.py --steps 1 --warmup 1 --bodies 200000 --cpu-sample 2000 --skip-e2e > gpurun_out/$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
