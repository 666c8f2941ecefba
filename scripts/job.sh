#!/bin/bash
# scratch GPU job: glyph-composed text; warm runs compose their sprites anew
python -m pytest tests/test_gpu_api.py tests/test_gpu_png.py -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("png", d["png_stage"]["device_figures_per_s"], d["png_stage"]["phases_s"])
a=d["api_e2e"]
for k in ("cold","warm","warm_other"):
    print(k, round(a[k]["seconds"],3), a[k]["pngs"], a[k]["errors"], a[k]["png_mb"], {x: y for x, y in a[k]["phases_s"].items()})
PY
tail -3 gpurun_out/bench.err | cut -c1-300
