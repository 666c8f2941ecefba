#!/bin/bash
# scratch GPU job: what the driver runs at round end, on one GPU
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","scaling","vs_baseline","dtype","gpu_launches")})
print("roofline", d["roofline"]); print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
print("api", d["api_e2e"]["value"], d["api_e2e"]["warm"]["seconds"], d["api_e2e"]["warm"]["phases_s"])
print("png", d["png_stage"]["device_figures_per_s"], "parity", d["parity_checked"]["ok"])
PY
