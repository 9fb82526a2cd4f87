#!/bin/bash
# Development helper (GPU box): time every library under build/variants/ on the bench workload.
# usage: bash scripts/variants_run.sh <tag> [sweep.py args...]
tag=$1; shift
mkdir -p gpurun_out
for lib in build/variants/lib_*.so; do
  name=$(basename $lib .so); name=${name#lib_}
  echo "== $name" >> gpurun_out/${tag}.log
  TSQB_LIBRARY=$PWD/$lib timeout 300 python scripts/sweep.py "$@" >> gpurun_out/${tag}.log 2>&1
done
cat gpurun_out/${tag}.log
