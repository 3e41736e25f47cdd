#!/bin/bash
# ncu launch list of the reference CUDA step (oracle/_ref) on C2: 1 warm-up + 2 timed steps.
# Usage (GPU box): bash scripts/ref_ncu.sh  -> gpurun_out/ref_launches.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ref_launches.csv \
    python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/ref_ncu.log 2>&1
