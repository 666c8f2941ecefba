#!/bin/bash
# scratch GPU job: 2 GPUs -- the multi-rank parity test at HEAD, then the bench at N=2
nvidia-smi -L
python -m pytest tests/test_gpu_api.py -m gpu -x -q -k two_gpu 2>&1 | tail -15 > gpurun_out/pytest_2gpu.log
cat gpurun_out/pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
print("collective", json.dumps(d["collective"]))
print("e2e", json.dumps(d["e2e"])[:700])
print("parity", json.dumps(d["parity_checked"])[:900])
PY
tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
