#!/bin/bash
# scratch GPU job: what the driver runs at round end, on one GPU
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","scaling","vs_baseline","dtype","gpu_launches")})
print("roofline", d["roofline"]["frac"], d["step_roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
a=d["api_e2e"]; print("api", a["value"], "cold", a["cold"]["seconds"], "warm", a["warm"]["seconds"], a["warm_other"]["seconds"], a["warm"]["phases_s"])
print("png", d["png_stage"]["device_figures_per_s"], d["png_stage"]["device_ratio"], "parity", d["parity_checked"]["ok"], d["stage_ms"])
PY
