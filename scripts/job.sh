#!/bin/bash
# GPU job run by scripts/gpusnap.sh (rewritten per call while developing); this version: what the driver runs at
# round end on one GPU -- the GPU tests, the smoke run, the default bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "api", d["api_e2e"]["value"], d["api_e2e"]["warm"]["seconds"], d["api_e2e"]["cold"]["seconds"], d["api_e2e"]["warm"]["pngs"], d["api_e2e"]["warm"]["errors"], "png", d["png_stage"]["device_figures_per_s"], "parity", d["parity_checked"]["ok"], d["clocks"]["reasons"])
PY
