#!/bin/bash
# scratch GPU job: host profile of the timed step loop (1 GPU)
python bench.py --steps 200 --warmup 5 --no-e2e --no-png --no-api-e2e --no-verify --no-cpu-baseline --profile-host gpurun_out/host_prof > gpurun_out/bench.json 2> gpurun_out/bench.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], d["stage_ms"])
PY
head -70 gpurun_out/host_prof.rank0 | cut -c1-200
