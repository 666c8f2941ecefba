#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
Q="--steps 3 --warmup 3 --no-e2e --no-png --no-cpu-baseline --no-api-e2e --no-verify"
python bench.py $Q > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:region_stats_kernel|rasterise_kernel' -c 3 -o gpurun_out/prof_k2a_k3 python bench.py $Q > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log | cut -c1-300
ls -la gpurun_out/
