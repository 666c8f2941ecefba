"""N > 1 host logic on CPU: two gloo ranks resolve the pooled extrema of a shared orbit sequence."""

import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])


def _run(world):
    port = _free_port()
    procs = [
        subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_worker.py"), str(r), str(world), port],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
        for r in range(world)
    ]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, out[-2000:] + err[-4000:]
    assert f"GLOO_OK world={world}" in outs[0][0]


def test_prefix_percentiles_two_gloo_ranks():
    _run(2)


def test_prefix_percentiles_three_gloo_ranks_uneven_blocks():
    _run(3)
