#!/bin/bash
# scratch GPU job (rewritten per gpurun call): strong-scaling point N=1 of config 4 (1000 orbits on one GPU)
timeout 1500 python bench.py --total-orbits 1000 --steps 10 --warmup 3 --no-e2e --no-png --no-api-e2e --no-cpu-baseline --verify-orbits 2 > gpurun_out/strong1.json 2> gpurun_out/strong1.err
echo "rc=$?"; cut -c1-1200 gpurun_out/strong1.json; tail -3 gpurun_out/strong1.err | cut -c1-300
nvidia-smi --query-gpu=memory.used --format=csv,noheader
