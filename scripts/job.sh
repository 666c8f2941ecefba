#!/bin/bash
# scratch GPU job: deferred K4 on a companion context
python -m pytest tests/test_gpu_api.py tests/test_gpu_png.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --no-e2e --no-png > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
a=d["api_e2e"]
for k in ("cold","warm","warm_other"):
    print(k, round(a[k]["seconds"],3), a[k]["pngs"], a[k]["errors"], a[k]["png_mb"], a[k]["phases_s"])
PY
tail -3 gpurun_out/bench.err | cut -c1-300
