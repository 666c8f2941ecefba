#!/bin/bash
# scratch GPU job (8 GPUs): the default bench line exactly as the driver launches it
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","n_gpus","ms_per_step","scaling","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["frac_of_h2d_ceiling"], "parity", d["parity_checked"]["ok"])
a=d["api_e2e"]; print("api", a["value"], a["n_gpus"], "cold", a["cold"]["seconds"], "warm", a["warm"]["seconds"], a["warm"]["pngs"], a["warm"]["errors"], a["warm"]["phases_s"])
print(d["collective"]["wait_us"], d["clocks"], d["stage_ms"])
PY
tail -3 gpurun_out/bench8.err | cut -c1-300
