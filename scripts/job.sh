#!/bin/bash
# scratch GPU job: round-2 ncu evidence (launch list of the timed steps; full captures of K1 / K2a / K3 / K4)
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-api-e2e --no-verify --no-cpu-baseline --png-orbits 2"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:collapse_stream|region_stats_kernel|rasterise_kernel" -c 6 -o gpurun_out/r2_prof_step $CMD > gpurun_out/ncu_step.log 2>&1
echo "step capture rc=$?"; tail -2 gpurun_out/ncu_step.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k "regex:png_encode_kernel" -s 1 -c 2 -o gpurun_out/r2_prof_png $CMD > gpurun_out/ncu_png.log 2>&1
echo "png capture rc=$?"; tail -2 gpurun_out/ncu_png.log | cut -c1-200
ls -la gpurun_out/
