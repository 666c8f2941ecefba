#!/bin/bash
# scratch GPU job (2 GPUs): multi-rank parity test + the default bench line at N=2, as the driver launches it
python -m pytest tests/test_gpu_api.py -m gpu -x -q -k two_gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","n_gpus","ms_per_step","scaling","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "parity", d["parity_checked"]["ok"])
a=d["api_e2e"]; print("api", a["value"], a["n_gpus"], a["warm"]["seconds"], a["warm"]["pngs"], a["warm"]["errors"], a["warm"]["results"], a["warm"]["phases_s"])
print(d["collective"]["wait_us"], d["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/ref2.json 2> gpurun_out/ref2.err
echo "ref2 rc=$?"; cut -c1-300 gpurun_out/ref2.json
tail -3 gpurun_out/bench2.err | cut -c1-300
