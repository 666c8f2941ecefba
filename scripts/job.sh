#!/bin/bash
# scratch GPU job: K4 adaptive match window, with / without per-row tile lists
python -m pytest tests/test_gpu_png.py -m gpu -x -q 2>&1 | tail -2
for rt in 1 0; do
CSG_PNG_ROW_TILES=$rt python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --no-e2e > gpurun_out/bench_rt$rt.json 2> gpurun_out/bench.err
python - $rt <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_rt{sys.argv[1]}.json").read().strip().splitlines()[-1])
p=d["png_stage"]; print("row_tiles", sys.argv[1], "png figs/s", round(p["device_figures_per_s"]), "ratio", round(p["device_ratio"],2), {k: round(v,4) for k,v in p["phases_s"].items()})
a=d["api_e2e"]; print("   api", round(a["value"],2), round(a["warm"]["seconds"],3), "png_mb", a["warm"]["png_mb"], {k: v for k,v in a["warm"]["phases_s"].items() if k.startswith("png") or k=="figures_host"})
PY
done
