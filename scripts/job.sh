#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
python -m pytest tests/test_gpu_png.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_png_api.log
cat gpurun_out/pytest_png_api.log
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_png.py --deselect tests/test_gpu_api.py 2>&1 | tail -8 > gpurun_out/pytest_rest.log
cat gpurun_out/pytest_rest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
print("png_stage", json.dumps(d["png_stage"])[:1500])
print("api_e2e", json.dumps(d["api_e2e"])[:2500])
print("e2e", json.dumps(d["e2e"])[:800])
print("parity", d["parity_checked"]["ok"], d["parity_checked"]["failures"])
PY
tail -3 gpurun_out/bench.err | cut -c1-300
mkdir -p gpurun_out/pngs; python - <<'PY'
# one sample figure of the directory driver for eyeballing
import os, sys, numpy as np, tempfile, shutil
sys.path.insert(0, os.getcwd())
from tests.test_gpu_api import _write_tree, _run_driver
import pathlib
d = pathlib.Path(tempfile.mkdtemp())
_write_tree(d); os.chdir(d)
res = _run_driver()
pngs = sorted(str(p) for p in pathlib.Path("FAST_plots").rglob("*.png"))
print(len(pngs), pngs[:2])
for p in pngs[:2] + pngs[-1:]:
    shutil.copy(p, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "gpu_snap/r2f/gpurun_out/pngs", os.path.basename(p)))
PY
ls -la gpurun_out/pngs
