#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
python -m pytest tests/test_gpu_png.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_png_api.log
cat gpurun_out/pytest_png_api.log
python bench.py --steps 10 --warmup 3 --api-ref > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
print("png_stage", json.dumps(d["png_stage"])[:1500])
print("api_e2e", json.dumps(d["api_e2e"])[:3000])
print("cpu", json.dumps(d["cpu_baseline"])[:600])
print("parity", d["parity_checked"]["ok"], d["parity_checked"]["failures"])
PY
tail -3 gpurun_out/bench.err | cut -c1-300
