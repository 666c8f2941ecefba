#!/usr/bin/env python
"""Time K1 (collapse) alone on a synthetic FAST shard for several slab-kernel configurations.

usage (on a GPU box): python scripts/k1_sweep.py [orbits]   -> one line per configuration
Each configuration runs in a fresh subprocess (the overrides are read from the environment).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from configurable_spectrograms_b200 import _lib
from configurable_spectrograms_b200.engine import Batch
n_orb = int(sys.argv[1])
dev = torch.device("cuda", 0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
ctx = _lib.Context(0, stream=s.cuda_stream)
P, E = 64, 96
Ts = [800, 903, 800, 903] * n_orb
total = sum(T * P * E for T in Ts)
cubes = torch.poisson(torch.full((total,), 2.0, device=dev))
cubes[torch.rand(total, device=dev) < 0.01] = float("nan")
bits = np.zeros(P, np.uint8)
pa = (np.arange(P) + 0.5) * 360.0 / P
for g, ranges in enumerate(([(0, 360)], [(0, 30), (330, 360)], [(150, 210)], [(40, 140), (210, 330)])):
    m = np.zeros(P, bool)
    for lo, hi in ranges: m |= (pa >= lo) & (pa <= hi)
    bits |= (m.astype(np.uint8) << g)
b = Batch(ctx, np.float32, n_groups=4)
off = 0
for T in Ts:
    b.add_file(None, bits, shape=(T, P, E), device_ptr=cubes.data_ptr() + 4 * off); off += T * P * E
for _ in range(3): b.collapse()
ms = []
for _ in range(7):
    ctx.timer_start(0); b.collapse(); ctx.timer_stop(0); ms.append(ctx.timer_ms(0))
ms = float(np.median(ms))
gb = (4 * total + sum(5 * T * E * 4 for T in Ts)) / 1e9
print(json.dumps({"ms": ms, "GBps": gb / ms * 1e3, "kernels": [str(k) for k in b.d_files]}))
""" % ROOT

def run(env_extra, n_orb):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD, str(n_orb)], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(json.dumps(env_extra), line, flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    configs = [{}]
    for c in configs:
        run(c, n)
