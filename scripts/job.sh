#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
CSG_API_PROFILE=$PWD/gpurun_out/api_profile.txt python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("png_stage", json.dumps(d["png_stage"])[:900])
print("api_e2e", json.dumps(d["api_e2e"])[:4200])
print("parity", d["parity_checked"]["ok"], d["parity_checked"]["failures"])
PY
tail -2 gpurun_out/bench.err | cut -c1-200
python scripts/bench_config5.py > gpurun_out/config5.json 2> gpurun_out/config5.err; cat gpurun_out/config5.json | cut -c1-2500; tail -3 gpurun_out/config5.err
