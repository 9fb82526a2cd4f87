#!/bin/bash
# Development helper (GPU box): the other BASELINE.json configurations, device-resident, one GPU.
# usage: bash scripts/sweep_configs.sh <tag>
tag=${1:-cfg}
mkdir -p gpurun_out
G=$((1<<30))
{
for b in 4096 16384 65536 262144 1048576 4194304; do timeout 300 python scripts/sweep.py text $G $b; done
timeout 400 python scripts/sweep.py random $((4*G)) 262144
timeout 300 python scripts/sweep.py rep8 $((2*G)) 1048576
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
