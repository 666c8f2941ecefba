#!/bin/bash
# scratch GPU job (rewritten per gpurun call)
nproc; free -g | head -2; df -h /dev/shm /tmp | cat
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; cat gpurun_out/pytest.log
