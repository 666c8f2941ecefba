#!/bin/bash
# scratch GPU job (rewritten per gpurun call): the multi-rank parity worker on 8 ranks (6 orbits: two empty shards)
CSG_TEST_WORLD=8 timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k two_gpu 2>&1 | tail -30 > gpurun_out/worker8.log
tail -30 gpurun_out/worker8.log | cut -c1-400
grep -a "rank[0-9]\]:" gpurun_out/multigpu_worker_failure.log 2>/dev/null | tail -30
exit 0
