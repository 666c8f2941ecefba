#!/bin/bash
# scratch GPU job: K2a loads-in-flight variants (A/B through CSG_LIBRARY); layout change validation
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_png.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
for v in head f2_b6 f3_b6 f4_b5 f4_b4; do
  CSG_LIBRARY=$PWD/variants/libcsgpu_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify --no-e2e --no-png --no-api-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - $v <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "region_stats", round(d["stage_ms"]["region_stats"],4), "step", round(d["ms_per_step"],4), "fallbacks", d["stage_ms"]["percentile_regions_needing_radix_fallback"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
