#!/bin/bash
# scratch GPU job: all-thread cProfile of one steady-state directory call
CSG_API_PROFILE=$PWD/gpurun_out/api_profile.txt python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --no-png > gpurun_out/bench.json 2> gpurun_out/bench.err
grep -v "^$" gpurun_out/api_profile.txt | awk '/Ordered by: internal time/{p=1} p' | cut -c1-150 | head -50
