#!/bin/bash
# scratch GPU job: K4 kernel capture after the dirty-range / adaptive-window changes; one sample figure
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-api-e2e --no-verify --no-cpu-baseline --png-orbits 2 --profile-region png"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:png_encode" -c 2 -o gpurun_out/r2_prof_png_after $CMD > gpurun_out/ncu_png.log 2>&1
echo "png capture rc=$?"; tail -2 gpurun_out/ncu_png.log | cut -c1-200
python - <<'PY'
import os, sys, glob, shutil, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from configurable_spectrograms_b200.fast.batch_directory import FAST_plot_spectrograms_directory
args = type("A", (), {"seed": 4})()
work = tempfile.mkdtemp(prefix="sample_", dir="/dev/shm")
cwd = os.getcwd()
try:
    from oracle import ref_driver as RD
    RD.prepare_directory(work, 3, 4)
    os.chdir(work)
    res = FAST_plot_spectrograms_directory("./FAST_data", output_base="./FAST_plots/", y_scale="linear", z_scale="log", colormap="turbo",
                                           max_processing_percentile=99, max_workers=4, progress_json_path="./progress.json", verbose=False)
    pngs = sorted(glob.glob("./FAST_plots/**/*.png", recursive=True))
    print(len(pngs), "pngs", [os.path.basename(p) for p in pngs[:3]])
    for p in pngs[:2]:
        shutil.copy(p, os.path.join(cwd, "gpurun_out", "sample_" + os.path.basename(p)))
finally:
    os.chdir(cwd); shutil.rmtree(work, ignore_errors=True)
PY
ls -la gpurun_out | head
