#!/bin/bash
# scratch GPU job: block cache skips fenced blocks; full GPU tests + bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "parity", d["parity_checked"]["ok"], "png", d["png_stage"]["device_figures_per_s"])
a=d["api_e2e"]
for k in ("cold","warm","warm_other"):
    print(k, round(a[k]["seconds"],3), a[k]["pngs"], a[k]["errors"], {x: y for x, y in a[k]["phases_s"].items() if not x.startswith("png/")})
PY
tail -3 gpurun_out/bench.err | cut -c1-300
