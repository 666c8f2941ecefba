#!/bin/bash
# scratch GPU job (4 GPUs): weak + strong series at N=4 and N=2
run() { # name nproc devices args...
  name=$1; n=$2; dev=$3; shift 3
  CUDA_VISIBLE_DEVICES=$dev timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
  python - $name <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("  ", d["n_gpus"], d["scaling"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "parity", (d.get("parity_checked") or {}).get("ok"), "wait", d["collective"].get("wait_us"))
except Exception as e:
    print("   no line:", e)
PY
}
run weak4 4 0,1,2,3 --steps 20 --warmup 5 --no-png --no-api-e2e
run strong4 4 0,1,2,3 --steps 20 --warmup 5 --total-orbits 1000 --no-png --no-api-e2e --no-e2e --no-verify
run weak2 2 0,1 --steps 20 --warmup 5 --no-png --no-api-e2e
run strong2 2 0,1 --steps 20 --warmup 5 --total-orbits 1000 --no-png --no-api-e2e --no-e2e --no-verify
