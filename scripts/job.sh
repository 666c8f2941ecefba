#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 --api-ref > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 4500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
