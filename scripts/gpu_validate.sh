#!/bin/bash
# One GPU-box call: parity tests, both bench arms, ncu launch list and one --set full capture of the two hot kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_validate.sh v7'
tag=${1:-vX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
cat gpurun_out/${tag}_bench_ours.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'encode_batch_kernel|decode_split_kernel' -c 2 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 0 --no-cpu --e2e-steps 0 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
