#!/bin/bash
# scratch GPU job: planning threads 16 (default) / 4 / 1
for t in 16 4 1; do
  CSG_RENDER_THREADS=$t python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-verify --no-e2e --no-png > gpurun_out/bench_t$t.json 2> gpurun_out/bench_t$t.err
  python - $t <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_t{sys.argv[1]}.json").read().strip().splitlines()[-1])
a=d["api_e2e"]; print("threads", sys.argv[1], "warm", round(a["warm"]["seconds"],3), round(a["warm_other"]["seconds"],3), "figures_host", a["warm"]["phases_s"]["figures_host"], a["warm_other"]["phases_s"]["figures_host"], "tile", a["warm"]["phases_s"]["png/tile_tables"])
PY
done
