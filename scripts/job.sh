#!/bin/bash
# scratch GPU job: K4 warps per block 4 (default) / 2 / 1
for v in default w2 w1; do
  lib=$PWD/configurable_spectrograms_b200/libcsgpu.so; [ $v != default ] && lib=$PWD/variants/libcsgpu_$v.so
  CSG_LIBRARY=$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-verify --no-e2e --no-api-e2e --png-orbits 16 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - $v <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    p=d["png_stage"]; print(sys.argv[1], "figs/s", round(p["device_figures_per_s"]), "encode", round(p["phases_s"]["encode_kernel_and_sizes"],4), "figures", p["figures"], "ratio", round(p["device_ratio"],2))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
CSG_LIBRARY=$PWD/variants/libcsgpu_w1.so python -m pytest tests/test_gpu_png.py -m gpu -x -q 2>&1 | tail -2
